"""Import shim: lets the UNMODIFIED reference import `dm_env` in this image.

TEST INFRASTRUCTURE ONLY (used by oracle/reference_runner.py).  dm_env carries
no arithmetic of the hot path; the stand-in is bsuite_b200/dm_env_compat.py,
loaded by file path so that importing the shim never imports the engine.
"""
import importlib.util as _util
import os as _os
import sys as _sys

_NAME = 'bsuite_b200.dm_env_compat'
if _NAME in _sys.modules:
  _compat = _sys.modules[_NAME]
else:
  _path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..', '..', '..', 'bsuite_b200',
                        'dm_env_compat.py')
  _spec = _util.spec_from_file_location(_NAME, _os.path.normpath(_path))
  _compat = _util.module_from_spec(_spec)
  _sys.modules[_NAME] = _compat
  _spec.loader.exec_module(_compat)

Environment = _compat.Environment
StepType = _compat.StepType
TimeStep = _compat.TimeStep
restart = _compat.restart
transition = _compat.transition
termination = _compat.termination
truncation = _compat.truncation
specs = _compat.specs
