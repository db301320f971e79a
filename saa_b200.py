"""Import alias: `import saa_b200` loads the package that lives in the directory
`synchronization-avoiding-algorithms_b200/` (its name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synchronization-avoiding-algorithms_b200")
_spec = importlib.util.spec_from_file_location(
    "saa_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["saa_b200"] = _mod
_spec.loader.exec_module(_mod)
