"""Import shim: the package directory is `auto-dynamic-deeplab_b200/` (not a valid Python
identifier), so `import add_b200` loads it under this name."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "auto-dynamic-deeplab_b200"
_spec = importlib.util.spec_from_file_location(
    "add_b200", _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["add_b200"] = _mod
_spec.loader.exec_module(_mod)
