"""Import shim: the package directory is `multi-h_b200/` (hyphenated, as the project is named); this module makes it
importable as `multih_b200` by loading that directory as a package and replacing itself in sys.modules."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "multi-h_b200")
_spec = _u.spec_from_file_location("multih_b200", _os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = _u.module_from_spec(_spec)
_sys.modules["multih_b200"] = _mod
_spec.loader.exec_module(_mod)
