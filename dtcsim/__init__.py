"""Import alias: ``dtcsim`` resolves to the package directory
``noise-resilience-in-discrete-time-crystal-realizations-on-quantum-computers_b200/`` (whose
name is not a valid Python identifier).  All code lives there."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "noise-resilience-in-discrete-time-crystal-realizations-on-quantum-computers_b200")
__path__.insert(0, _PKG_DIR)
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
