"""The C-ABI library loads on a CPU-only box, exports every symbol include/gkr_b200.h declares, and
fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from gkr_b200 import _lib
from tests.conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gkr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(gkr_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_exported():
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/gkr_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared


def test_version_string():
    assert b"sm_100a" in _lib.lib().gkr_version()


@pytest.mark.skipif(has_gpu(), reason="checks the no-device error path")
def test_no_device_is_a_loud_error():
    L = _lib.lib()
    ctx = C.c_void_p()
    rc = L.gkr_ctx_create(0, C.byref(ctx))
    assert rc == -2 and not ctx.value
    assert b"no CPU fallback" in L.gkr_last_error()
    from gkr_b200 import Prover
    with pytest.raises(_lib.GkrError):
        Prover(0)


def test_communicator_entry_points_reject_bad_arguments_without_a_device():
    """argument checks come before any CUDA call: GKR_ERR_INVALID, a message, nothing aborts"""
    L = _lib.lib()
    assert L.gkr_comm_init_shared(None, 2, 0, b"/gkr_test") == -1
    assert b"gkr_comm_init_shared" in L.gkr_last_error()
    assert L.gkr_comm_init(None, 2, 0, None) == -1
    ctxs = (C.c_void_p * 3)()
    devs = (C.c_int32 * 3)(0, 0, 0)
    assert L.gkr_comm_create(3, devs, ctxs) == -1           # not a power of two
    assert all(not c for c in ctxs)


def test_product_does_not_import_oracle():
    """the product package must never route through oracle/"""
    pkg = os.path.join(ROOT, "gkr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "imports oracle/", "").replace("import oracle/", ""), f


def test_public_header_is_plain_c():
    """the drop-in boundary is a C ABI: include/gkr_b200.h must compile as C99 (and as C++) without extensions"""
    import os
    import shutil
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "hc.c")
        with open(src, "w") as f:
            f.write('#include "gkr_b200.h"\nint main(void) { gkr_fr x; (void)x; return GKR_OK; }\n')
        inc = os.path.join(root, "include")
        subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", src], check=True)
        subprocess.run([cc, "-x", "c++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", src], check=True)
