"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Build the C oracle (and, when /root/reference is present, oracle/_ref) once per session."""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)
    yield


@pytest.fixture(scope="session")
def coop_emu(tmp_path_factory):
    """tests/coop_emu/coop_emu.cpp built for the host: the four-lane element solve with four threads playing the lanes."""
    import ctypes as C
    import subprocess
    out = str(tmp_path_factory.mktemp("coop_emu") / "libcoop_emu.so")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Werror", "-shared", "-fPIC", "-pthread",
           os.path.join(ROOT, "tests", "coop_emu", "coop_emu.cpp"), "-o", out]
    c = subprocess.run(cmd, capture_output=True, text=True)
    assert c.returncode == 0, c.stderr
    lib = C.CDLL(out)
    vp, f32, u32 = C.c_void_p, C.c_float, C.c_uint32
    lib.coop_emu_sweep.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, f32, f32, f32, f32, vp, vp, vp, u32]
    lib.coop_emu_sweep.restype = C.c_int
    return lib
