"""Import shim: the product package lives in `implicitbvh.jl_b200/` (a directory name Python cannot
import directly because of the dot). `import ibvh_b200` loads that package under this module name."""
import importlib.util as _u
import os as _os
import sys as _sys

_pkg = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "implicitbvh.jl_b200")
_spec = _u.spec_from_file_location(__name__, _os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = _u.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
