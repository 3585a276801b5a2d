"""B200-native polar-contour hot path."""
__version__ = "0.1.0"
