"""Import the reference as installed by baseline/install_reference.py (baseline/_ref) and apply, at run time and
without touching its files, the patches it needs to run at all (all probed):
  * matplotlib is not installed: a stub package goes on sys.path (only when the real one is missing);
  * `class Segment(Detect)` inherits the wrong head in the snapshot (nn/modules/head.py:436: `Detect` emits nc+64
    channels); the parent that matches `Segment` is `polarpaperDetect` (nn/modules/head.py:364-433).
Test / benchmark infrastructure: the product package never imports this."""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "ultralytics"))


def _matplotlib_stub():
    try:
        import matplotlib  # noqa: F401
        return
    except ImportError:
        pass
    stub = tempfile.mkdtemp(prefix="ycr_mpl_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    with open(os.path.join(stub, "matplotlib", "__init__.py"), "w") as f:
        f.write("def use(*a, **k):\n    pass\ndef rc(*a, **k):\n    pass\n"
                "class _F:\n    def __getattr__(self, n):\n        return _F()\n"
                "    def __call__(self, *a, **k):\n        return _F()\n"
                "font_manager = _F()\nrcParams = {}\n")
    for sub in ("pyplot", "image", "colors", "figure", "patches"):   # what utils/plotting.py and the callbacks import
        with open(os.path.join(stub, "matplotlib", sub + ".py"), "w") as f:
            f.write("def __getattr__(n):\n    def f(*a, **k):\n        return None\n    return f\n")
    sys.path.insert(0, stub)


_loaded = None


def load():
    """-> the imported `ultralytics` package of baseline/_ref with the two run-time patches applied."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_reference.py` in the build container")
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="ycr_cfg_"))
    _matplotlib_stub()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import ultralytics
    import ultralytics.nn.modules.head as rhead
    rhead.Segment.__bases__ = (rhead.polarpaperDetect,)
    rhead.Detect.forward = rhead.polarpaperDetect.forward   # Segment.__init__ captures Detect.forward (head.py:443)
    _loaded = ultralytics
    return ultralytics


def model_yaml() -> str:
    """The polar segmentation model definition of the reference's repository root (nm=36, nc=10)."""
    return os.path.join(REF_DIR, "yolov8-seg.yaml")


def reference_criterion(nc, rays, strides, box=7.5, cls=0.5, device="cpu"):
    """The reference's v8SegmentationLoss (utils/loss.py:772) on the CPU, built on a stub of the two objects its
    constructor reads (`model.args`, `model.model[-1]`), so the loss path can be driven without the conv backbone."""
    from types import SimpleNamespace
    import torch
    load()
    import ultralytics.utils.loss as rloss

    class _Head(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.nc, self.nm, self.no, self.reg_max = nc, rays, nc + rays, 16
            self.stride = torch.tensor(strides, dtype=torch.float32)
            self.w = torch.nn.Parameter(torch.zeros(1))

    class _Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.ModuleList([_Head()])
            self.args = SimpleNamespace(box=box, cls=cls, dfl=1.5, overlap_mask=True)

    return rloss.v8SegmentationLoss(_Model().to(device))
