"""The C++ host-side mirror (rust-snappy_b200/host/snap.hpp) over the C ABI."""
import os
import subprocess

import pytest

from conftest import DATA, ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def _build():
    import __graft_entry__ as g
    g.build_cuda()
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    hdr = os.path.join(ROOT, "rust-snappy_b200", "host", "snap.hpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-o", EXE, src, "-L" + os.path.join(ROOT, "rust-snappy_b200"),
                               "-lsnapb200", "-Wl,-rpath," + os.path.join(ROOT, "rust-snappy_b200")])


def test_cpp_mirror_builds_and_refuses_without_gpu():
    import torch
    _build()
    rc = subprocess.run([EXE, os.path.join(DATA, "html")], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert rc.returncode == 0, rc.stdout + rc.stderr
    else:
        assert rc.returncode == 3, rc.stdout + rc.stderr      # NoDevice: there is no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["html", "urls.10K", "fireworks.jpeg", "alice29.txt"])
def test_cpp_mirror_roundtrips_on_gpu(name):
    _build()
    rc = subprocess.run([EXE, os.path.join(DATA, name)], capture_output=True, text=True, timeout=300)
    assert rc.returncode == 0, rc.stdout + rc.stderr
