"""Import alias: the package directory is `pcc-rl_b200/` (not a valid Python identifier), so
`import pcc_rl_b200` loads that directory as the package `pcc_rl_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pcc-rl_b200")
_spec = _u.spec_from_file_location("pcc_rl_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["pcc_rl_b200"] = _mod
_spec.loader.exec_module(_mod)
