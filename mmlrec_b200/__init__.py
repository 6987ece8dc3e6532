"""Importable alias for the product package.

The package directory is named after the reference repository
(``mmlrec-a-unified-multi-task-and-multi-scenario-learning-benchmark-for-recommendation_b200``),
which is not a valid Python identifier; this shim makes it importable as ``mmlrec_b200`` by
pointing the package search path at that directory and running its ``__init__``.
"""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))
PACKAGE_DIR = _os.path.join(
    _os.path.dirname(_HERE),
    "mmlrec-a-unified-multi-task-and-multi-scenario-learning-benchmark-for-recommendation_b200")
__path__ = [PACKAGE_DIR]
with open(_os.path.join(PACKAGE_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(PACKAGE_DIR, "__init__.py"), "exec"))
