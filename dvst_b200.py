"""Import shim: the package directory is named `dino-video-summarization-transformer_b200/` (not a valid Python
identifier), so `import dvst_b200` loads it from that directory under the importable name `dvst_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dino-video-summarization-transformer_b200")
_spec = importlib.util.spec_from_file_location(
    "dvst_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dvst_b200"] = _mod
_spec.loader.exec_module(_mod)
