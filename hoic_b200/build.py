"""Build libhoic_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per kernel of the env-step pipeline (each has its own compact shared-memory layout, chosen by
H_KERNEL before csrc/dev_model.h is included) plus the C ABI; compiled in parallel, linked into one shared library."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# (source, object, extra flags): k_p3.cu and k_a.cu are compiled twice, once per capacity tier
UNITS = [("hoic_capi.cu", "hoic_capi.o", []), ("k_p1.cu", "k_p1.o", []), ("k_p1.cu", "k_p1w.o", ["-DP1_WIDE"]), ("k_p2.cu", "k_p2.o", []), ("k_p3.cu", "k_p3.o", []),
         ("k_p3.cu", "k_p3s.o", ["-DP3_SMALL"]), ("k_p3.cu", "k_p3sb.o", ["-DP3_SMALL", "-DP3_BUDGET"]), ("k_p3.cu", "k_p3c.o", ["-DP3_SMALL", "-DP3_CONT"]), ("k_p3.cu", "k_p3fc.o", ["-DP3_CONT"]), ("k_tail.cu", "k_tail.o", []), ("k_glue.cu", "k_glue.o", []),
         ("k_p12.cu", "k_p12.o", []), ("k_a.cu", "k_a.o", []), ("k_a.cu", "k_as.o", ["-DP3_SMALL"])]
OUT = os.environ.get("HOIC_BUILD_OUT", os.path.join(HERE, "libhoic_b200.so"))      # HOIC_BUILD_OUT: A/B variant builds
OBJDIR = os.path.join(HERE, "build") if "HOIC_BUILD_OUT" not in os.environ else OUT + ".obj"
DEPS = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".log")] + [
    os.path.join(HERE, "..", "include", "hoic.h"), os.path.join(HERE, "..", "include", "hoic_blob.h"), os.path.abspath(__file__)]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--fmad=true", "-prec-div=false", "-prec-sqrt=false", "-ftz=true", "-Xptxas", "-v"]
EXTRA = os.environ.get("HOIC_NVCC_FLAGS", "").split()     # e.g. -DP3_WPS=12 for A/B builds


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def _compile(u):
    unit, objname, extra = u
    obj = os.path.join(OBJDIR, objname)
    r = subprocess.run([NVCC] + FLAGS + EXTRA + extra + ["-c", "-o", obj, os.path.join(CSRC, unit)], capture_output=True, text=True)
    return objname, obj, r


def build(force=False, verbose=False):
    if not force and not needs_build() and not EXTRA:
        return OUT
    os.makedirs(OBJDIR, exist_ok=True)
    with ThreadPoolExecutor(len(UNITS)) as ex:
        results = list(ex.map(_compile, UNITS))
    log = ""
    for unit, obj, r in results:
        log += f"==== {unit}\n{r.stderr}"
        if verbose or r.returncode != 0:
            sys.stderr.write(f"==== {unit}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {unit}")
    r = subprocess.run([NVCC, "-shared", "-o", OUT] + [obj for _, obj, _ in results], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
        f.write(log)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
