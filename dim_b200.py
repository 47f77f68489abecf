"""Import alias: `import dim_b200` loads the package that lives in ./dyadic-interaction-modeling_b200/
(the directory name required by the repo layout is not a valid Python identifier)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dyadic-interaction-modeling_b200")
_spec = _u.spec_from_file_location("dim_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["dim_b200"] = _mod
_spec.loader.exec_module(_mod)
