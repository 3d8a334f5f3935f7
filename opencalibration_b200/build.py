"""In-tree build of libocb.so (CUDA, sm_100a) and libocb_host.so (C++ mirror of the reference entry points).

nvcc cross-compiles without a GPU; the built .so files are git-ignored but travel to the GPU box.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall", "-shared"]
CUDA_SOURCES = ["hamming_top2.cu", "hamming_lists.cu", "score_models.cu", "fit_models.cu", "link_tail.cu", "hamming_tensor.cu", "pipe_probe.cu", "ocb_capi.cu"]
HOST_SOURCES = ["linalg.cpp", "models.cpp", "homography_decompose.cpp", "distort_keypoints.cpp", "link_batch.cpp", "partition.cpp", "match_features.cpp", "guided_match.cpp", "ransac.cpp", "relax_refit.cpp", "graph_wire.cpp", "graph_wire_capi.cpp", "flat_shim.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and r.stdout:
        print(r.stdout)


def build_cuda(force=False, verbose=False):
    out = os.path.join(PKG, "libocb.so")
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, "ocb_internal.cuh"), os.path.join(ROOT, "include", "ocb.h")]
    if force or _newer(out, deps):
        _run(["nvcc"] + NVCC_FLAGS + ["-o", out] + srcs, verbose)
    return out


def build_host(force=False, verbose=False):
    out = os.path.join(PKG, "libocb_host.so")
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES if os.path.exists(os.path.join(HOST, s))]
    if not srcs:
        return None
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")] + \
        [os.path.join(HOST, "pow10_table.inc"), os.path.join(ROOT, "include", "ocb.h"),
         os.path.join(ROOT, "include", "ocb_wire.h")]
    if force or _newer(out, deps):
        _run(["g++", "-std=c++17", "-O3", "-fPIC", "-Wall", "-Wextra", "-ffp-contract=off", "-fopenmp", "-shared",
              "-I", os.path.join(ROOT, "include"), "-I", HOST, "-o", out] + srcs +
             ["-L", PKG, "-locb", "-Wl,-rpath,$ORIGIN"], verbose)
    return out


def build_all(force=False, verbose=False):
    a = build_cuda(force, verbose)
    b = build_host(force, verbose)
    return a, b


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
