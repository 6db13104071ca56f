"""TEST INFRASTRUCTURE: import the UNMODIFIED reference (/root/reference/tornadox) on top of the NumPy-backed
``jax`` shim in oracle/jaxshim.  Only usable in the build container (the reference tree does not travel)."""

import os
import pathlib
import sys
import types

SHIM = pathlib.Path(__file__).resolve().parent / "jaxshim"
REFERENCE = pathlib.Path(os.environ.get("TORNADOX_REFERENCE", "/root/reference"))


def available():
    return (REFERENCE / "tornadox" / "__init__.py").exists()


def import_reference():
    """Returns the reference ``tornadox`` package running on the shim."""
    if not available():
        raise RuntimeError(f"{REFERENCE} is not present (the reference only exists in the build container)")
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__file__", "").startswith(str(SHIM)):
        raise RuntimeError("a real jax is already imported; the shim is not needed")
    for path in (str(SHIM), str(REFERENCE)):
        if path not in sys.path:
            sys.path.insert(0, path)
    # tornadox/__init__.py:10 imports a setuptools_scm-generated file that is not in the tree
    ver = types.ModuleType("tornadox._version")
    ver.version = "0+reference.shim"
    sys.modules.setdefault("tornadox._version", ver)
    import tornadox

    return tornadox


if __name__ == "__main__":
    tdx = import_reference()
    print("reference imported from", tdx.__file__)
