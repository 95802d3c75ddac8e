"""Builds the C++ host mirror (librec_b200/host/*.cpp) into _lib/liblibrec_b200_host.so, linked
against the CUDA library liblibrec_b200.so that sits next to it."""
import os
import subprocess

from .. import _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_build.LIB_DIR, "liblibrec_b200_host.so")
_SRCS = [os.path.join(_HERE, f) for f in ("librec_host.cpp", "librec_host_capi.cpp")]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _SRCS + [os.path.join(_HERE, "librec_host.hpp"), os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "librec_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    _build.build()
    if not force and not is_stale():
        return LIB_PATH
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-Wextra", "-o", LIB_PATH] + _SRCS + \
          ["-L" + _build.LIB_DIR, "-llibrec_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed:\n" + r.stdout[-6000:])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True))
