"""Build the C-ABI shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m fs_uae_image_enhancer_project_b200.build [--force] [-v]

Every translation unit is recompiled only when one of the files it includes (transitively, inside csrc/ and include/)
is newer than its object; objects live in build/<hash of the flags>/ so that A/B builds (FSUAE_EXTRA_NVCC_FLAGS +
FSUAE_LIB_PATH) keep their own.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.environ.get("FSUAE_LIB_PATH") or os.path.join(PKG_DIR, "libfsuae_enhancer.so")   # override: A/B builds side by side
# approximate transcendentals + flush-to-zero: tensor-core builds only, never the fp32 build
TC = ["--use_fast_math"]
# (source, object, extra flags): the tensor-core units are compiled twice -- bf16 operands and, with -DFSUAE_OPERAND_FP16, fp16 operands
UNITS = [("abi.cu", "abi.o", []), ("fp32_path.cu", "fp32_path.o", []), ("synth.cu", "synth.o", []),
         ("bf16_tc.cu", "bf16_tc.o", TC), ("bf16_tc.cu", "fp16_tc.o", TC + ["-DFSUAE_OPERAND_FP16"]),
         ("mega.cu", "bf16_mega.o", TC), ("mega.cu", "fp16_mega.o", TC + ["-DFSUAE_OPERAND_FP16"])]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]
_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(path: str, seen: set[str] | None = None) -> set[str]:
    """`path` and every file it includes with quotes, transitively (csrc/ and include/ only)."""
    seen = set() if seen is None else seen
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path, encoding="utf-8") as fh:
        for name in _INC.findall(fh.read()):
            for base in (os.path.dirname(path), CSRC, INCLUDE):
                cand = os.path.join(base, name)
                if os.path.exists(cand):
                    _deps(cand, seen)
                    break
    return seen


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libfsuae_enhancer.so (units whose object is newer than all their sources are kept)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("FSUAE_EXTRA_NVCC_FLAGS", "").split()      # debugging aids (e.g. -DFSUAE_EPI_TIMING)
    tag = hashlib.sha1(" ".join(NVCC_FLAGS + extra).encode()).hexdigest()[:8]
    objdir = os.path.join(PKG_DIR, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src, objname, unit_flags in UNITS:
        obj = os.path.join(objdir, objname)
        objs.append(obj)
        newest = max(os.path.getmtime(p) for p in _deps(os.path.join(CSRC, src)))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            continue
        cmd = [nvcc, *extra, *NVCC_FLAGS, *unit_flags, "-I", INCLUDE, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((objname, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = None
    for objname, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = failed or objname
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    if procs or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs):
        subprocess.run([nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
