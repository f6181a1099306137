"""Import shim: registers `physics-based-climate-model_b200/` (not a valid Python identifier) as
the package `pcm_b200`.  `import pcm_b200` from the repo root is the supported entry."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "physics-based-climate-model_b200")
_spec = importlib.util.spec_from_file_location("pcm_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pcm_b200"] = _mod
_spec.loader.exec_module(_mod)
