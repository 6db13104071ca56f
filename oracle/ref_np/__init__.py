"""ORACLE (test infrastructure): NumPy fp64 restatement of the tornadox hot path.  See oracle/__init__.py."""
from . import driver, filters, prior, problems  # noqa: F401
