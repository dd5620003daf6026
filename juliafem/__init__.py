"""Import shim: makes the on-disk package directory `juliafem.jl_b200/` importable as
`juliafem.jl_b200` (a directory name containing a dot cannot be imported directly)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "juliafem.jl_b200")
if "juliafem.jl_b200" not in _sys.modules:
    _spec = _u.spec_from_file_location("juliafem.jl_b200", _os.path.join(_dir, "__init__.py"),
                                       submodule_search_locations=[_dir])
    _mod = _u.module_from_spec(_spec)
    _sys.modules["juliafem.jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
    jl_b200 = _mod
else:
    jl_b200 = _sys.modules["juliafem.jl_b200"]
