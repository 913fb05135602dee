"""Build libboda_b200.so in-tree with nvcc for sm_100a (no torch involvement; plain C ABI)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libboda_b200.so")
SOURCES = ["b200_compute.cu", "b200_conv_fwd.cu", "caffe_prototxt.cu", "wisdom.cu", "b200_shard.cu", "b200_abi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "-ccbin", "g++", "--expt-relaxed-constexpr", "-Xptxas", "-v" if os.environ.get("B200_PTXAS_V") else "-warn-spills"]


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [os.path.join(HERE, "..", "include", "boda_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    objs, jobs = [], []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, src):
            jobs.append([NVCC] + FLAGS + ["-c", src, "-o", obj])
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for " + cmd[-3])
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(OUT):
        # static cudart: the .so has no libcuda/libcudart link-time dependency (driver entry points are resolved at run time)
        run([NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xcompiler", "-fPIC", "-ccbin", "g++"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
