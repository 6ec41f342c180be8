"""Builds libuavsim.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

Three translation units:
  csrc/uavsim_f64.cu  parity kernels      -fmad=false (no FMA contraction; see uavsim_kernels.cuh)
  csrc/uavsim_f32.cu  throughput kernels  --use_fast_math (FMA contraction, flush-to-zero, approximate div/sqrt)
  csrc/uavsim_api.cu  C-ABI + host logic
nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libuavsim.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", str(ROOT / "include"), "-I", str(CSRC)]
UNITS = {
    "uavsim_f64.cu": ["-fmad=false"],
    "uavsim_f32.cu": ["--use_fast_math"],   # flush-to-zero MUFU sequences, approximate div/sqrt: tolerance build
    "uavsim_api.cu": [],
}


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libuavsim.so cannot be built (there is no CPU fallback)")
    return exe


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "uavsim.h", Path(__file__)]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Builds (or reuses) the library.  Safe under `torch.distributed.run`, where every rank imports the package at
    once: one process builds under an exclusive file lock, the others wait for it and find the fingerprint fresh; the
    library is linked under a temporary name and renamed into place, so nobody can dlopen a half-written file."""
    import fcntl
    stamp = OBJ / "fingerprint"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    OBJ.mkdir(exist_ok=True)
    with open(OBJ / "lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
                return LIB                      # another rank built it while this one waited
            return _build_locked(fp, stamp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(fp: str, stamp: Path, verbose: bool) -> Path:
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in UNITS.items():
        obj = OBJ / (src[:-3] + ".o")
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose and out:
            print(out)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    tmp = LIB.with_name(LIB.name + ".tmp%d" % os.getpid())
    link = [nvcc, *ARCH, "-shared", "-o", str(tmp), *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB)
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p, os.path.getsize(p), "bytes")
