"""Builds test-only native helpers into tests/native/_build/."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")


def build_host_numerics():
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libhost_numerics.so")
    srcs = [os.path.join(HERE, "host_numerics.cpp"),
            os.path.join(HERE, "..", "..", "better-binary-quantization_b200", "csrc", "bbq_numerics.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-x", "c++", srcs[0], "-o", so])
    return so
