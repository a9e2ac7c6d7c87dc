"""Builds ``_gpfx_tf_ops.so`` when TensorFlow is importable; a no-op (returns None) otherwise."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent


def build_tf_shim():
    try:
        import tensorflow as tf
    except ImportError:
        return None
    lib = HERE.parent / "lib" / "libgpflux_b200.so"
    out = HERE / "_gpfx_tf_ops.so"
    cmd = ["g++", "-std=c++17", "-shared", "-fPIC", "-O2", str(HERE / "gpfx_tf_ops.cc"), "-o", str(out),
           *tf.sysconfig.get_compile_flags(), *tf.sysconfig.get_link_flags(),
           "-DGOOGLE_CUDA=1", "-I/usr/local/cuda/include", str(lib), f"-Wl,-rpath,{lib.parent}"]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build_tf_shim())
