"""Import shim: `import ycr_b200` loads the package that lives in the directory
`yolo-contour-regression_b200/` (a name the layout contract fixes but Python cannot import
directly because of the hyphens).  After import, `ycr_b200.<submodule>` resolves inside that
directory like any regular package."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "yolo-contour-regression_b200")
_spec = importlib.util.spec_from_file_location(
    "ycr_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ycr_b200"] = _mod
_spec.loader.exec_module(_mod)
