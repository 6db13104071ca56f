"""pytest plugin (TEST INFRASTRUCTURE): run the reference's own test-suite on the NumPy-backed jax shim.

    python -m pytest -p oracle.shim_pytest_plugin -p no:cacheprovider /root/reference/tests/test_ek1.py
"""
from oracle.run_reference import import_reference

import_reference()
