"""Build the product library llama.swift_b200/libb200llama.so IN-TREE with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  No torch, no CPU fallback:
csrc/engine.cu (kernels + C ABI, nvcc), csrc/host_math.cpp (host libm constants) and csrc/host_text.cpp (tokenizer /
sampler drop-ins), both g++.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libb200llama.so")
SRCS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
SRCS.append(os.path.join(HERE, "..", "include", "b200_llama.h"))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in SRCS)


def build(force: bool = False, verbose: bool = False, defs=(), out: str = LIB) -> str:
    """defs/out: development A/B builds (e.g. defs=["-DB200_ROWLOOP=0"], out=.../libb200llama_r0.so, picked up through
    the B200_LIB environment variable); the product is the default build."""
    if not force and not defs and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_host = os.path.join(HERE, "csrc", "host_math.o")
    # host constants: plain g++ semantics (no contraction, F16C for the fp16 conversions)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mf16c", "-c",
                           os.path.join(HERE, "csrc", "host_math.cpp"), "-o", obj_host])
    # host-side tokenizer / sampler drop-ins (same libstdc++ / libm calls as the reference's utils.cpp)
    obj_text = os.path.join(HERE, "csrc", "host_text.o")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-c",
                           os.path.join(HERE, "csrc", "host_text.cpp"), "-o", obj_text])
    # the token loop of -[LlamaPredictOperation main] above the C ABI
    obj_run = os.path.join(HERE, "csrc", "host_runner.o")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-c",
                           os.path.join(HERE, "csrc", "host_runner.cpp"), "-o", obj_run])
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-fmad=false",                       # FMAs only where the reference has them (explicit fmaf / fma.rn.f32x2)
           "-Xcompiler", "-fPIC", "-shared", "-ccbin", "g++",
           os.path.join(HERE, "csrc", "engine.cu"), obj_host, obj_text, obj_run, "-o", out, "-lcudart"] + list(defs)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defs=defs, out=outs[0] if outs else LIB))
