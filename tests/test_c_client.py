"""The C ABI used from plain C (gcc, no Python in the call path): builds tests/c_abi_client.c
against include/*.h and the in-tree libegpu_alloc.so — what a cgo binding links against."""
import hashlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "elastic-gpu-agent_b200", "lib")
EXE = os.path.join(LIBDIR, "c_abi_client")


def build():
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi_client.c"), "-L", LIBDIR, "-legpu_alloc", "-Wl,-rpath," + LIBDIR, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_client_builds_and_fails_loudly_without_a_gpu(egpu):
    import torch
    build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 77 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_client_on_gpu(egpu):
    build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c abi client ok" in r.stdout
    assert "hash(3-07) = " + hashlib.sha256(b"3-07").hexdigest()[:8] in r.stdout
