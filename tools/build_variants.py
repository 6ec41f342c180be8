#!/usr/bin/env python
"""Developer tool: build A/B variants of the throughput kernels (csrc/uavsim_f32.cu with extra -D flags) as
<package>/variants/libuavsim_<tag>.so, reusing the product build's other objects.  A variant is selected at
run time with UAVSIM_LIB=<path> (see _native.py).  Usage:

    python tools/build_variants.py tag1:-DUAVSIM_W2=7,-DUAVSIM_B2=3 tag2:-DFOO ...
"""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import PKG_DIR  # noqa: E402

sys.path.insert(0, str(PKG_DIR))
import build as B  # noqa: E402


def one(spec):
    tag, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    out = PKG_DIR / "variants"
    out.mkdir(exist_ok=True)
    obj = out / ("uavsim_f32_%s.o" % tag)
    cmd = [B._nvcc(), *B.ARCH, *B.COMMON, *B.UNITS["uavsim_f32.cu"], "-Xptxas=-v", *flags, "-c", str(B.CSRC / "uavsim_f32.cu"), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return tag, "FAILED\n" + r.stdout + r.stderr
    lib = out / ("libuavsim_%s.so" % tag)
    link = [B._nvcc(), *B.ARCH, "-shared", "-o", str(lib), str(obj), str(B.OBJ / "uavsim_f64.o"), str(B.OBJ / "uavsim_api.o")]
    r2 = subprocess.run(link, capture_output=True, text=True)
    if r2.returncode != 0:
        return tag, "LINK FAILED\n" + r2.stdout + r2.stderr
    info = []
    lines = (r.stdout + r.stderr).splitlines()
    for i, ln in enumerate(lines):
        if "Compiling entry function" in ln and "step_kernelIfLb0ELi" in ln and "ELi3ELb0ELi1" in ln:
            info.append(ln.split("'")[1][24:44] + " | " + lines[i + 2].strip() + " | " + lines[i + 3].strip())
    obj.unlink()
    return tag, "\n".join(info)


if __name__ == "__main__":
    B.build()
    with ThreadPoolExecutor(8) as ex:
        for tag, msg in ex.map(one, sys.argv[1:]):
            print("==", tag)
            print(msg)
