"""Importable alias of the package directory ``dgvit-depth-goal-guided-vision-transformer-_b200``
(whose name is not a valid Python identifier)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "dgvit-depth-goal-guided-vision-transformer-_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
