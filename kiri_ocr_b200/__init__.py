"""Import shim: the package source lives in ``kiri-ocr_b200/`` (the layout the build
contract names); a hyphen is not importable, so this module re-points ``__path__`` there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "kiri-ocr_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
