"""Import alias: `better-binary-quantization_b200/` (the package directory the repo layout prescribes) is not
an importable name, so `import bbq_b200` re-exports it.  No code lives here."""
import importlib.util as _u
import os as _os
import sys as _sys

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "better-binary-quantization_b200")
_name = "better_binary_quantization_b200"
if _name not in _sys.modules:
    _spec = _u.spec_from_file_location(_name, _os.path.join(_impl, "__init__.py"), submodule_search_locations=[_impl])
    _mod = _u.module_from_spec(_spec)
    _sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_mod = _sys.modules[_name]
globals().update({k: getattr(_mod, k) for k in _mod.__all__})
_native = _mod._native
__all__ = list(_mod.__all__)
