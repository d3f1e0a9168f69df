"""Import shim: the product package lives in ``i-dqn_b200/`` (the name the project layout asks for, which
is not a valid Python identifier); this makes it importable as ``idqn_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "i-dqn_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
