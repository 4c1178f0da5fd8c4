"""Import alias: `import wseg_b200` loads the package directory `1-stage-wseg_b200/`
(whose name is not a valid Python identifier) and registers it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "1-stage-wseg_b200")
_spec = importlib.util.spec_from_file_location("wseg_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["wseg_b200"] = _mod
_spec.loader.exec_module(_mod)
