"""CPU-only: the C-ABI library loads, exports every symbol include/vsb200.h declares,
the ctypes stub covers exactly those symbols, and the product fails loudly (no CPU
fallback, no oracle import) when no GPU is present."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from volatilespace_b200 import build
    build.build()
    return ctypes.CDLL(build.LIB)


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), "libvsb200.so does not export " + s


def test_ctypes_stub_matches_header():
    from volatilespace_b200 import _lib
    assert sorted(list(_lib.SIGNATURES) + ["vsb_last_error", "vsb_sym_schedule"]) == declared_symbols()
    _lib.load()
    # the SASS post-pass either ran or the library says why not (build.status() is the same string)
    from volatilespace_b200 import build
    assert _lib.sym_schedule().split(" ")[0] in ("resched", "ptxas")
    assert _lib.sym_schedule().split(" ")[0] == build.status()["sym_sched"]


def test_no_compute_without_gpu_is_loud(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from volatilespace_b200._lib import VsbError
    from volatilespace_b200.device import Device
    with pytest.raises(VsbError):
        Device(0)
    from volatilespace_b200.phys_editor import Physics
    with pytest.raises(VsbError):
        Physics(0)


def test_product_never_imports_the_oracle():
    """volatilespace_b200/ must not reference oracle/ (a product path through the oracle
    would void every parity claim)."""
    pkg = os.path.join(ROOT, "volatilespace_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "vs_oracle" not in text and "import oracle" not in text, f
                assert "/root/reference" not in text, f
    code = ("import sys; import volatilespace_b200.phys_editor, volatilespace_b200.phys_body, "
            "volatilespace_b200.phys_vessel, volatilespace_b200.dist; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_sass_uses_tma_bulk_copies():
    """The pair kernels stage their source tiles with 1-D TMA bulk copies (UBLKCP)."""
    from volatilespace_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", build.LIB], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UBLKCP" in out.stdout and "MUFU.RSQ64H" in out.stdout and "DFMA" in out.stdout


def test_pinned_buffers_degrade_to_plain_arrays_without_a_gpu(lib):
    """Page-locking the shim's result arrays is an optimisation: without a device (or when the
    registration is refused) pinned_empty returns an ordinary writable array of the asked shape and
    leaves no pending CUDA error behind."""
    import numpy as np
    from volatilespace_b200.device import pinned_empty
    for shape, dt in (((100000, 2), np.float64), (70000, np.int64), ((3, 5), np.float64), (0, np.float64)):
        a = pinned_empty(shape, dt)
        assert a.shape == (shape if isinstance(shape, tuple) else (shape,)) and a.dtype == dt
        assert a.flags.c_contiguous and a.flags.writeable and a.ctypes.data % 8 == 0
        a[...] = 7
        assert (a == 7).all()
