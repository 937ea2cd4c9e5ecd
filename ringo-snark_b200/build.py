"""Builds libringo_b200.so (sm_100a only) in-tree with nvcc.  No torch involved.

    python ringo-snark_b200/build.py [--force]

Objects go to ringo-snark_b200/build/, the library to ringo-snark_b200/libringo_b200.so
(git-ignored, but it travels to the GPU box with the snapshot).
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libringo_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-ccbin", "/usr/bin/g++"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False, defines=(), tag=""):
    """defines/tag: build an experimental variant (tools/) into libringo_b200<tag>.so."""
    global OBJ, LIB
    if tag:
        OBJ = os.path.join(HERE, "build" + tag)
        LIB = os.path.join(HERE, f"libringo_b200{tag}.so")
    FLAGS.extend("-D" + d for d in defines)
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            jobs.append([NVCC] + ARCH + FLAGS + ["-c", s, "-o", o])
    # host-only sources (the csprng restatement): g++, no FMA contraction (Go on amd64 never fuses a*b+c)
    for s in sorted(glob.glob(os.path.join(CSRC, "*.cpp"))):
        o = os.path.join(OBJ, os.path.basename(s)[:-4] + ".host.o")
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            jobs.append(["/usr/bin/g++", "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-pthread"]
                        + ["-D" + d for d in defines] + ["-c", s, "-o", o])
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode:
                    sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
                if res.returncode:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or not _newer(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart", "-ldl", "-Xcompiler", "-pthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    tag = next((a[6:] for a in sys.argv[1:] if a.startswith("--tag=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, tag=tag))
