"""Importable alias of the ``point-gnn_b200/`` package directory.

The product lives in ``/root/repo/point-gnn_b200`` (the layout the project brief
names); a hyphen is not legal in a Python module name, so this stub package
points its ``__path__`` there and runs that directory's ``__init__.py``.
``import pointgnn_b200.models.gnn`` therefore loads
``point-gnn_b200/models/gnn.py``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'point-gnn_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _os, _f, _real
