"""Import shim: the product package lives in the directory literally named `interfaceadvection.jl_b200/`
(the name the build contract asks for).  A dot is not legal inside a Python package name, so this tiny
parent package makes `import interfaceadvection.jl_b200` resolve to that directory."""
import importlib.util as _u
import os as _os
import sys as _sys

_root = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "interfaceadvection.jl_b200")
_name = __name__ + ".jl_b200"
if _name not in _sys.modules:
    _spec = _u.spec_from_file_location(_name, _os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
    _mod = _u.module_from_spec(_spec)
    _sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = _sys.modules[_name]
