"""In-tree nvcc build of the C-ABI library (sm_100a only).

``python -m qsparse_b200.build`` compiles every ``csrc/*.cu`` with
``nvcc -gencode arch=compute_100a,code=sm_100a`` and links
``qsparse_b200/lib/libqsparse_b200.so``.  nvcc cross-compiles without a GPU, so
this runs in the CPU build container; the resulting ``.so`` is git-ignored but
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
OBJ_DIR = PKG / "build"
LIB_PATH = LIB_DIR / "libqsparse_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # every ATen op in the reference rounds separately: never contract a*b+c
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math",
    "--expt-relaxed-constexpr",
] + os.environ.get("QSB_EXTRA_NVCC_FLAGS", "").split()   # development only (e.g. -DQSB_SELECT_TIMING)


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stamp(src: Path, deps: list[Path]) -> str:
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in [src, *deps]:
        h.update(p.read_bytes())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> Path:
    nvcc = find_nvcc()
    LIB_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h"))
    sources = sorted(p for p in CSRC.glob("*.cu") if not p.name.startswith("bench_"))

    def compile_one(src: Path):
        obj = OBJ_DIR / (src.stem + ".o")
        stamp_file = OBJ_DIR / (src.stem + ".stamp")
        stamp = _stamp(src, headers)
        if not force and obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
            return obj, False, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        stamp_file.write_text(stamp)
        return obj, True, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    rebuilt = any(r[1] for r in results)
    if verbose:
        for _, _, log in results:
            if log:
                print(log)
    # relink also when an earlier run compiled objects but died before linking (e.g. a closed pipe)
    stale = LIB_PATH.exists() and any(r[0].stat().st_mtime > LIB_PATH.stat().st_mtime for r in results)
    if rebuilt or stale or not LIB_PATH.exists() or force:
        cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *[str(r[0]) for r in results],
               "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(path)
