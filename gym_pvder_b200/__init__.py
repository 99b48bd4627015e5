"""Importable alias of the package directory ``gym-solarpvder-environment_b200/`` (a hyphen is
not a legal Python identifier).  All code lives there; this module only points ``__path__`` at it
and runs its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "gym-solarpvder-environment_b200")
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
with open(_init) as _fh:
    exec(compile(_fh.read(), _init, "exec"))
del _os, _fh, _init
