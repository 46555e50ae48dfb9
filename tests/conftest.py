import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
HOSTEMU_SO = os.path.join(ROOT, "tests", "hostemu", "libtcb200_hostemu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def O():
    import oracle
    oracle.lib()
    return oracle


def build_hostemu():
    """Test-only host build of the device task code (see tests/hostemu/hostemu.cpp)."""
    src = os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")
    deps = [src] + [os.path.join(ROOT, "threshold_crypto_b200", "csrc", f) for f in ("fp.cuh", "tower.cuh", "scheme.cuh")]
    if os.path.exists(HOSTEMU_SO) and all(os.path.getmtime(HOSTEMU_SO) >= os.path.getmtime(d) for d in deps):
        return HOSTEMU_SO
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", HOSTEMU_SO])
    return HOSTEMU_SO


@pytest.fixture(scope="session")
def emu():
    from threshold_crypto_b200._lib import Engine
    return Engine(build_hostemu())


@pytest.fixture(scope="session")
def gpu_engine():
    from threshold_crypto_b200._lib import Engine
    return Engine()   # raises loudly if libtcb200.so or the GPU is missing


def fr_bytes(vals):
    return np.frombuffer(b"".join((int(v) % R).to_bytes(32, "little") for v in vals), np.uint8).copy()


def rand_fr(rng, n):
    return fr_bytes([int.from_bytes(rng.bytes(40), "little") for _ in range(n)])


def hx(s):
    return np.frombuffer(bytes.fromhex(s), np.uint8).copy()


def hxs(lst):
    return np.stack([hx(s) for s in lst]) if lst else np.zeros((0, 0), np.uint8)
