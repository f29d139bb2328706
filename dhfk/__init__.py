"""Import alias.  The product package lives in
``dh-aug-dh-forward-kinematics-model-driven-augmentation-for-3d-human-pose-estimation_b200/``;
that directory name is not a Python identifier, so ``import dhfk`` maps onto it: this module
adopts the real directory as its ``__path__`` and executes its ``__init__.py``.  Sub-modules
(``dhfk.functional``, ``dhfk.camera`` ...) therefore load from the real directory.
"""
import os as _os

_REAL = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "dh-aug-dh-forward-kinematics-model-driven-augmentation-for-3d-human-pose-estimation_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _f
