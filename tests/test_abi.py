"""CPU: the CUDA library builds, loads and exports every symbol include/tcb200.h declares
(no compute calls — there is no GPU here), and refuses to initialise without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "threshold_crypto_b200", "csrc", "libtcb200.so")


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tcb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(SO), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(SO)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"missing export {s}"


def test_hostemu_exports_the_host_buffer_surface():
    from conftest import build_hostemu
    lib = ctypes.CDLL(build_hostemu())
    for s in declared_symbols():
        if s.endswith("_dev") or s.startswith(("tcb_probe", "tcb_selftest", "tcb_miller")):
            continue
        assert hasattr(lib, s), s


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to start (tcb_init != 0)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from threshold_crypto_b200._lib import Engine, TcbError
    with pytest.raises(TcbError):
        Engine()


def test_sass_is_sm100a_integer_code():
    """The shipped cubin targets sm_100a; the Fp multiply is carry-chained IMAD.WIDE (IMAD.WIDE.U32.X with predicate carries),
    the shared-memory Miller kernel stages its operands with 128-bit LDS/STS and takes its inputs through the TMA unit (UBLKCP),
    and there is no tensor-core instruction anywhere (integer modular arithmetic)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", SO], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    syms = subprocess.run(["cuobjdump", "-symbols", SO], capture_output=True, text=True).stdout.split()
    fn = next(w for w in syms if w.startswith("_Z13k_miller_quad"))
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, SO], capture_output=True, text=True).stdout
    assert "Function : " + fn in sass
    assert sass.count("IMAD.WIDE.U32.X") > 300, "carry-chained wide multiply-accumulates missing"
    assert "UBLKCP" in sass, "bulk asynchronous (TMA) input staging missing"
    assert sass.count("LDS.128") > 50 and sass.count("STS.128") > 50, "shared-memory operand staging missing"
    for tensor_op in ("HMMA", "IMMA", "UTCHMMA", "UTCIMMA"):
        assert tensor_op not in sass
