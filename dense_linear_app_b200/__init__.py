"""Importable alias of the package directory ``dense-linear-app_b200/`` (a hyphen cannot be
imported).  The sources live there; this file only points ``__path__`` at them."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dense-linear-app_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
