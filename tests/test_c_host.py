"""The drop-in boundary is a C ABI: a strict C99 host (tests/c_host/host_only.c) is compiled with gcc against
include/xpbd_fem_b200.h, linked with the built library and run - no GPU needed (host-only scene)."""
import os
import subprocess

from __graft_entry__ import ROOT, build, load_package

build()
xf = load_package()


def test_c99_host_compiles_links_and_runs(tmp_path):
    exe = str(tmp_path / "host_only")
    lib_dir = os.path.dirname(xf.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_host", "host_only.c"), "-o", exe, "-L", lib_dir, "-lxpbd_fem_b200", "-Wl,-rpath," + lib_dir]
    c = subprocess.run(cmd, capture_output=True, text=True)
    assert c.returncode == 0, c.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    lines = r.stdout.splitlines()
    assert lines[0].startswith("verts 72 elements 180 colours 24 chained ")
    assert int(lines[0].split()[-1]) >= 500                  # per mille of corner uses the chained sweep keeps in a thread
    assert lines[1].startswith("substep rc %d: " % xf.XF_ERR_CUDA) and "no CPU compute path" in lines[1]
    assert lines[2] == "bad hint rc %d" % xf.XF_ERR_COLORING
