"""In-tree build of librobotgnn.so (nvcc, sm_100a only).  Used by __graft_entry__.build().
Each .cu is compiled to an object in parallel (only when stale), then linked into the shared library next to this file."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librobotgnn.so")
SOURCES = ["api.cu", "csr.cu", "aggregate.cu", "film_complete.cu", "film_ring.cu", "film_dense.cu", "film_mma.cu", "pose.cu", "edge_mlp.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_fused.cu", "gemm_dxfilm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(CSRC, "film_common.cuh"), os.path.join(HERE, "..", "include", "robot_gnn.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def needs_build():
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    return _stale(LIB, srcs + HEADERS)


def build(force=False, verbose=False):
    """Compile every CUDA source into multi-robot-perception-gnn-1_b200/librobotgnn.so."""
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        if force or _stale(obj, [src] + HEADERS):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        return job, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for (src, obj), r in ex.map(compile_one, jobs):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
            if verbose:
                sys.stderr.write(r.stderr)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if jobs or force or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking librobotgnn.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
