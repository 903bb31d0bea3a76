"""Import alias: the product lives in the directory `if-defense_b200/` (a name Python cannot import);
this shim makes it importable as `ifdefense_b200` without moving any file."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "if-defense_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
