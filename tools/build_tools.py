"""Build tools/_build/libb200seg_tools.so: experiments and diagnostics that are NOT part of the product library or its
public header (tcgen05.mma issue-rate probe, the tensor-core-depthwise variant of the fused inverted-residual kernel)
plus the standalone FHFMA throughput probe.    python tools/build_tools.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "team02-objectdetection_b200", "csrc")
OUT = os.path.join(HERE, "_build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"]


def build():
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "csrc", "probe.cu"), os.path.join(HERE, "csrc", "mbconv_tc.cu"), os.path.join(CSRC, "runtime.cu")]
    so = os.path.join(OUT, "libb200seg_tools.so")
    subprocess.check_call([NVCC, *ARCH, "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-I", CSRC, "-o", so, *srcs])
    exe = os.path.join(OUT, "fhfma_probe")
    subprocess.check_call([NVCC, *ARCH, "-o", exe, os.path.join(HERE, "fhfma_probe.cu")])
    return so, exe


if __name__ == "__main__":
    print(*build(), sep="\n")
