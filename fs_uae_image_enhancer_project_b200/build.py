"""Build the C-ABI shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m fs_uae_image_enhancer_project_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("FSUAE_LIB_PATH") or os.path.join(PKG_DIR, "libfsuae_enhancer.so")   # override: A/B builds side by side
# (source, object, extra flags): bf16_tc.cu is compiled twice -- bf16 operands and, with -DFSUAE_OPERAND_FP16, fp16 operands
UNITS = [("abi.cu", "abi.o", []), ("fp32_path.cu", "fp32_path.o", []), ("synth.cu", "synth.o", []),
         # approximate transcendentals + flush-to-zero: tensor-core builds only, never the fp32 build
         ("bf16_tc.cu", "bf16_tc.o", ["--use_fast_math"]),
         ("bf16_tc.cu", "fp16_tc.o", ["--use_fast_math", "-DFSUAE_OPERAND_FP16"])]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _newest_source_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "fsuae_enhancer.h")]
    return max(os.path.getmtime(p) for p in paths)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libfsuae_enhancer.so (skipped when the .so is newer than every source)."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(PKG_DIR, "build"), exist_ok=True)
    procs = []
    for src, objname, unit_flags in UNITS:
        obj = os.path.join(PKG_DIR, "build", objname)
        cmd = [nvcc, *NVCC_FLAGS, *unit_flags, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        for extra in os.environ.get("FSUAE_EXTRA_NVCC_FLAGS", "").split():   # debugging aids (e.g. -DFSUAE_EPI_TIMING)
            cmd.insert(1, extra)
        procs.append((objname, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
