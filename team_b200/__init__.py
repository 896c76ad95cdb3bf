"""Import alias: ``team_b200`` is the importable name of the package that lives in
``team-temporal-evolution-aware-multimodal-model_b200/`` (a directory name Python cannot
import directly).  All code is there; this file only extends ``__path__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "team-temporal-evolution-aware-multimodal-model_b200")
__path__.append(_real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
