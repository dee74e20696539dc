"""Import alias for the hyphenated package ``robotic-ultrasound-imaging_b200``.

``import rui_b200.env`` loads ``robotic-ultrasound-imaging_b200/env.py``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "robotic-ultrasound-imaging_b200")
__path__ = [_os.path.normpath(_real)]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _f.name, "exec"))
del _f, _real
