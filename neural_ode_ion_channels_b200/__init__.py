"""Importable alias of the ``neural-ode-ion-channels_b200/`` package directory.

The directory name the project layout prescribes contains hyphens and therefore cannot be
imported directly; this shim points ``__path__`` at it and runs its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'neural-ode-ion-channels_b200')
__path__ = [_real]
_init = _os.path.join(_real, '__init__.py')
with open(_init) as _fh:
    exec(compile(_fh.read(), _init, 'exec'))
del _fh, _init
