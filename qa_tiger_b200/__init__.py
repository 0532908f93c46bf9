"""Importable alias of the `qa-tiger_b200/` package (a hyphen is not a valid module name).

`import qa_tiger_b200` executes `qa-tiger_b200/__init__.py` with this module's __path__ pointing at
that directory, so `qa_tiger_b200.models.modules` etc. resolve to the files under `qa-tiger_b200/`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "qa-tiger_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
