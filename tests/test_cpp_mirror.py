"""The C++ mirror of the reference API (include/tcb200.hpp) running the reference's scheme tests
(tests/cpp/test_mirror.cpp): against the host-emulation build on the CPU box, against libtcb200.so on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp")


def _build_and_run(lib_path, exe):
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)
    subprocess.check_call(["g++", "-O1", "-std=c++17", SRC, "-o", exe, "-L" + libdir, "-l:" + libname, "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp mirror tests ok" in out.stdout


def test_cpp_mirror_on_host_emulation(tmp_path):
    from conftest import build_hostemu
    _build_and_run(build_hostemu(), str(tmp_path / "mirror_emu"))


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path):
    _build_and_run(os.path.join(ROOT, "threshold_crypto_b200", "csrc", "libtcb200.so"), str(tmp_path / "mirror_gpu"))
