"""CPU: the C-ABI shared library loads, exports every symbol include/tornadox_b200.h declares, and its
host-only entry points agree with the reference-derived golden vectors.  No compute calls (no GPU here)."""

import ctypes
import pathlib
import re

import numpy as np
import pytest

from helpers import assert_close, golden

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from tornadox_b200 import _lib, build

    build.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "tornadox_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tdx_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from tornadox_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes binding and header disagree"


def test_abi_version(lib):
    assert lib.tdx_abi_version() == 1


def test_struct_layouts_match_header():
    from tornadox_b200 import _lib

    assert ctypes.sizeof(_lib.ProblemDesc) == 4 + 4 + 4 * 8 + 4 + 4 + 8 * 8   # incl. 4 B padding before params
    assert ctypes.sizeof(_lib.StepArgs) == 4 * 8 + 4 + 4
    assert ctypes.sizeof(_lib.StepRule) == 4 + 4 + 6 * 8              # tdx_steprule
    assert ctypes.sizeof(_lib.FullStepResult) == 4 * 8 + 8            # tdx_full_step_result
    assert ctypes.sizeof(_lib.CtlState) == 3 * 8 + 2 * 8 + 4 * 4      # tdx_ctl_state
    # the C side agrees: a C translation unit that includes the header prints the same sizes
    import pathlib, shutil, subprocess, tempfile

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc:
        root = pathlib.Path(__file__).resolve().parent.parent
        with tempfile.TemporaryDirectory() as tmp:
            src = pathlib.Path(tmp) / "sizes.c"
            src.write_text('#include <stdio.h>\n#include "tornadox_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n", '
                           'sizeof(tdx_problem_desc), sizeof(tdx_step_args), sizeof(tdx_steprule), sizeof(tdx_full_step_result), '
                           'sizeof(tdx_ctl_state));return 0;}\n')
            exe = pathlib.Path(tmp) / "sizes"
            subprocess.run([cc, "-I", str(root / "include"), str(src), "-o", str(exe)], check=True)
            out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
        assert [int(x) for x in out] == [ctypes.sizeof(t) for t in (_lib.ProblemDesc, _lib.StepArgs, _lib.StepRule,
                                                                     _lib.FullStepResult, _lib.CtlState)]


@pytest.mark.parametrize("nu", [1, 2, 3, 4, 5])
def test_prior_and_preconditioner_match_reference(lib, nu):
    from tornadox_b200 import _lib

    g = golden()
    A, Ql = _lib.prior(nu)
    np.testing.assert_array_equal(A, g[f"prior/nu{nu}/A"])
    # built-in factor (extended precision) vs the reference's fp64 LAPACK dpotrf: cond(Q) * eps apart
    assert_close(Ql, g[f"prior/nu{nu}/Ql"], 1e-12 if nu < 5 else 1e-10, "Ql")
    # the factor the Python host hands to the library is the reference's, bit for bit
    np.testing.assert_array_equal(_lib.lapack_diffusion_sqrt(nu), g[f"prior/nu{nu}/Ql"])
    for dt in (0.1, 1e-3, 2.5):
        p, pinv = _lib.nordsieck(nu, dt)
        assert_close(p, g[f"prior/nu{nu}/p/{dt}"], 1e-14, "p")
        assert_close(pinv, g[f"prior/nu{nu}/pinv/{dt}"], 1e-14, "p_inv")


def test_invalid_arguments_are_reported(lib):
    from tornadox_b200 import _lib

    A = np.zeros((2, 2))
    pd = ctypes.POINTER(ctypes.c_double)
    assert lib.tdx_prior(0, A.ctypes.data_as(pd), A.ctypes.data_as(pd)) == 1
    assert b"nu" in lib.tdx_last_error()
    with pytest.raises(ValueError):
        _lib.prior(17)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the context cannot be created and the Python engine refuses to run."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = ctypes.c_void_p()
    rc = lib.tdx_create(ctypes.byref(handle), 0, 0)
    assert rc == 3 and b"no CPU fallback" in lib.tdx_last_error()
    from tornadox_b200 import _lib, ivp
    from tornadox_b200.engine import StepEngine

    with pytest.raises(_lib.TornadoxB200Error):
        StepEngine(ivp.vanderpol().spec, 4)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under tornadox_b200/ may reference it."""
    for path in (ROOT / "tornadox_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in re.sub(r"#.*", "", text), f"{path} mentions the oracle"


def test_c_client_links_and_runs_the_host_only_entry_points(lib):
    """A plain C program built against include/tornadox_b200.h and linked with the shared library (what a cgo / JNI / FFI
    binding does): the host-only entry points run without a GPU, tdx_create reports the missing device."""
    import pathlib, shutil, subprocess, tempfile

    from tornadox_b200 import _lib

    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    root = pathlib.Path(__file__).resolve().parent.parent
    libdir = pathlib.Path(_lib.LIB_PATH).resolve().parent
    program = r'''
#include <stdio.h>
#include "tornadox_b200.h"
int main(void) {
    double A[25], Ql[25], p[5], pinv[5];
    if (tdx_abi_version() != 1) return 10;
    if (tdx_prior(4, A, Ql) != TDX_OK) return 11;
    if (tdx_nordsieck(4, 0.1, p, pinv) != TDX_OK) return 12;
    if (tdx_prior(0, A, Ql) == TDX_OK || tdx_last_error()[0] == 0) return 13;
    printf("%.17g %.17g %.17g %.17g\n", A[1], Ql[0], p[0], pinv[4]);
    tdx_ctx* ctx = 0;
    int rc = tdx_create(&ctx, 0, TDX_F64);
    printf("%d\n", rc);
    if (rc == TDX_OK) tdx_destroy(ctx);
    return 0;
}
'''
    with tempfile.TemporaryDirectory() as tmp:
        src = pathlib.Path(tmp) / "client.c"
        src.write_text(program)
        exe = pathlib.Path(tmp) / "client"
        subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", str(root / "include"), str(src), "-o", str(exe),
                        "-L", str(libdir), "-ltornadox_b200", f"-Wl,-rpath,{libdir}"], check=True)
        out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    g = golden()
    assert float(out[0]) == g["prior/nu4/A"][0, 1] == 4.0
    assert abs(float(out[1]) - g["prior/nu4/Ql"][0, 0]) <= 1e-15
    assert abs(float(out[2]) - g["prior/nu4/p/0.1"][0]) <= 1e-14 * abs(g["prior/nu4/p/0.1"][0])
    assert abs(float(out[3]) - g["prior/nu4/pinv/0.1"][4]) <= 1e-14 * abs(g["prior/nu4/pinv/0.1"][4])
    import torch

    assert int(out[4]) == (0 if torch.cuda.is_available() else 3)
