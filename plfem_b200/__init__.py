"""Import shim: the product package lives in ``pl-fem-vectoriel_b200/`` (a
directory name Python cannot import directly because of the hyphens).  This
module makes it importable as ``plfem_b200`` by pointing ``__path__`` at that
directory and executing its ``__init__.py`` here."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "pl-fem-vectoriel_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
