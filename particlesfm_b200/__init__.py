"""Import shim: the package directory is named `particle-sfm_b200/` (not a valid Python
identifier); this package re-exports it as `particlesfm_b200` by pointing __path__ at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "particle-sfm_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
