"""Build libspsk.so (the C-ABI CUDA library, include/spsk.h) in-tree for sm_100a.

    python -m spsnet_b200.build [--force] [--verbose]

Plain nvcc, one object per .cu compiled in parallel, linked into spsnet_b200/_C/libspsk.so with the
static CUDA runtime (no torch headers, no torch symbols: the library is a pure C-ABI and is loaded
through ctypes by spsnet_b200/_lib.py, or by any host language's FFI).  Objects are rebuilt only when
their source (or a header) is newer.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT_DIR = ROOT / "_C"
OBJ_DIR = OUT_DIR / "obj"
LIB = OUT_DIR / "libspsk.so"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "spsk.h"]
    return max(h.stat().st_mtime for h in hs if h.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    hdr_m = _headers_mtime()
    jobs = []
    for src in sources:
        obj = OBJ_DIR / (src.stem + ".o")
        stale = force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m)
        if stale:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    failed = False
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(compile_one, jobs):
                if r.returncode != 0:
                    failed = True
                    sys.stderr.write(f"[spsk build] FAILED {src.name}\n{r.stdout}\n{r.stderr}\n")
                else:
                    (OBJ_DIR / (src.stem + ".ptxas.log")).write_text(r.stderr)
                    if verbose:
                        sys.stderr.write(f"[spsk build] {src.name}\n{r.stderr}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    objs = [OBJ_DIR / (s.stem + ".o") for s in sources]
    if jobs or force or not LIB.exists():
        # device-link is not needed (no relocatable device code); hide everything except extern "C" spsk_*
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
               "-cudart", "static", "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    lib = build(force=a.force, verbose=a.verbose)
    print(lib)


if __name__ == "__main__":
    main()
