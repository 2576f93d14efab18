"""Builds and loads tests/host_emul/emul.cpp: the kernels' per-problem C++
(blsq_core.cuh) compiled for the host behind the same C ABI.  TEST ONLY."""
import os
import subprocess

from bounded_lsq_b200._lib import Lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul.cpp")
OUT = os.path.join(HERE, "host_emul", "libblsq_hostemul.so")
_CSRC = os.path.join(os.path.dirname(HERE), "bounded_lsq_b200", "csrc")
CORE = os.path.join(_CSRC, "blsq_core.cuh")
TALL = [os.path.join(_CSRC, "blsq_tall_core.cuh"),
        os.path.join(_CSRC, "blsq_tall_common.cuh")]


def build():
    deps = [SRC, CORE] + TALL
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d)
                                   for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
           "-mfma", "-x", "c++", SRC, "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


class HostEmulLib(Lib):
    requires_cuda = False


_lib = None


def get():
    global _lib
    if _lib is None:
        _lib = HostEmulLib(build())
    return _lib
