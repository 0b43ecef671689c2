"""Import shim: makes ``import bundleadjustment.jl_b200`` resolve to the directory
``bundleadjustment.jl_b200/`` at the repository root (a dotted directory name cannot be imported
directly)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "bundleadjustment.jl_b200")
if "bundleadjustment.jl_b200" not in _sys.modules:
    _spec = _u.spec_from_file_location("bundleadjustment.jl_b200", _os.path.join(_dir, "__init__.py"),
                                       submodule_search_locations=[_dir])
    _mod = _u.module_from_spec(_spec)
    _sys.modules["bundleadjustment.jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = _sys.modules["bundleadjustment.jl_b200"]
