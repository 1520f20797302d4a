"""Import shim: the product package lives in the directory `splendor-rl-gym_b200/` (a name
Python cannot import directly); `import splendor_rl_gym_b200` loads it under this name."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / 'splendor-rl-gym_b200'
_spec = importlib.util.spec_from_file_location(__name__, _pkg_dir / '__init__.py',
                                               submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
