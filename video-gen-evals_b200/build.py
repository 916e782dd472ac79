"""Build libtag_b200.so (sm_100a only) in-tree with nvcc. No torch/pybind dependency: the library is
a plain CUDA-runtime shared object loaded through ctypes (see _lib.py)."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libtag_b200.so")
SOURCES = ["tag_api.cu", "k1_feature_fuse.cu", "k34_score.cu", "enc_kernels.cu", "gemm_simt.cu", "gemm_tc.cu", "tlayer_tc.cu", "tcn_block_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (needed to build libtag_b200.so for sm_100a)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """experiments=True builds libtag_b200_exp.so with -DTAG_EXPERIMENTS: the bottleneck switches (TAG_TC_DEBUG,
    TAG_TC_HALO, TAG_TC_PAIR, TAG_K1_DEBUG, TAG_FRAME_TABLE environment variables) exist only there, for tools/;
    the product library reads no environment variable."""
    OBJ = os.path.join(HERE, "build_exp" if experiments else "build")
    LIB = os.path.join(HERE, "libtag_b200_exp.so" if experiments else "libtag_b200.so")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "tag_b200.h"))
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + (["-DTAG_EXPERIMENTS"] + (["-DTAG_MBAR_NO_HINT"] if os.environ.get("TAG_BUILD_NO_HINT") else []) + (["-DTAG_EPI_LEADER_POLL"] if os.environ.get("TAG_BUILD_LEADER_POLL") else []) if experiments else [])

    def compile_one(src):
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src.replace(".cu", ".o"))
        stamp = op + ".sha"
        dg = _digest([sp] + headers) + ("exp" if experiments else "")
        if not force and os.path.exists(op) and os.path.exists(stamp) and open(stamp).read() == dg:
            return op, ""
        cmd = [nvcc] + flags + ["-c", sp, "-o", op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(dg)
        return op, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [r[0] for r in results]
    logs = "".join(r[1] for r in results)
    if verbose and logs:
        print(logs)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, experiments="--experiments" in sys.argv))
