"""Import shim: the package directory is named `tiled-mm_b200` (not a valid Python identifier), so
`import tiled_mm_b200` from the repo root resolves here and loads it under this name."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "tiled-mm_b200"
_spec = importlib.util.spec_from_file_location("tiled_mm_b200", _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tiled_mm_b200"] = _mod
_spec.loader.exec_module(_mod)
