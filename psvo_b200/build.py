"""Build libpsvo_b200.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m psvo_b200.build [--force] [--verbose]

Outputs psvo_b200/lib/libpsvo_b200.so (git-ignored, travels to the GPU box with the snapshot).
Objects are cached under build/ and rebuilt when a source or header is newer."""
import argparse
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# PSVO_BUILD_VARIANT=name + PSVO_BUILD_DEFS="-DX=1 ...": experimental build into lib/libpsvo_b200_<name>.so (own object cache),
# loaded with PSVO_LIB_PATH; the default build takes neither
VARIANT = os.environ.get("PSVO_BUILD_VARIANT", "")
OBJ = os.path.join(ROOT, "build", "obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "lib", "libpsvo_b200%s.so" % ("_" + VARIANT if VARIANT else ""))

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"] + (["-DPSVO_TC_TRACE"] if os.environ.get("PSVO_TC_TRACE") == "1" else []) + os.environ.get("PSVO_BUILD_DEFS", "").split()


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, obj, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (src, log))
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.abspath(__file__)]
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append((s, o))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for f in [ex.submit(_compile, s, o, verbose) for s, o in jobs]:
                f.result()
    if jobs or _newer(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                        "-Xcompiler", "-fPIC", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
