"""Import alias: the package directory is ``gr-fosphor_b200/`` (not a valid
Python identifier), so ``import gr_fosphor_b200`` extends its search path to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gr-fosphor_b200")
__path__.append(_real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
