"""In-tree native build: g++ for the host engine, nvcc (sm_100a only) for the device engine.

Both libraries are plain C-ABI shared objects loaded with ctypes (no torch headers, so a full
rebuild takes seconds and cross-compiles on a box without a GPU).  The .so files land in
``deeprec_b200/lib`` so they travel with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIB = os.path.join(ROOT, "lib")
OBJ = os.path.join(LIB, "obj")

HOST_SOURCES = ["host/host_engine.cc", "host/io_runtime.cc", "host/ssd_store.cc", "host/predict_codec.cc", "host/redis_store.cc", "host/tensor_pool.cc", "host/cpu_serving.cc", "host/ps_server.cc"]
CUDA_SOURCES = [
    "cuda/table_kernels.cu",
    "cuda/embedding_kernels.cu",
    "cuda/optimizer_kernels.cu",
    "cuda/dense_kernels.cu",
    "cuda/gemm_fp8.cu",
    "cuda/gemm_tcgen05.cu",
    "cuda/interaction_kernels.cu",
    "cuda/comm_kernels.cu",
    "cuda/sparse_pipeline.cu",
    "cuda/fused_interaction_gemm.cu",
    "cuda/tier_kernels.cu",
    "cuda/nvls.cu",
    "cuda/runtime.cu",
    "cuda/program_kernels.cu",
    "cuda/sparse_utils.cu",
    "cuda/serving_runtime.cu",
    "cuda/fused_ops.cu",
    "cuda/allocator.cu",
    "cuda/attention_kernels.cu",
]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _site_cutlass_include() -> list[str]:
    try:
        import flashinfer  # noqa: F401  (only for its vendored header tree)
        p = os.path.join(os.path.dirname(flashinfer.__file__), "data", "cutlass", "include")
        if os.path.isdir(p):
            return ["-I" + p]
    except Exception:
        pass
    return []


def _stamp(srcs: list[str], flags: list[str]) -> str:
    h = hashlib.sha1()
    for f in flags:
        h.update(f.encode())
    deps = list(srcs)
    for d in ("common", "cuda", "host"):
        dd = os.path.join(CSRC, d)
        if os.path.isdir(dd):
            for fn in sorted(os.listdir(dd)):
                if fn.endswith((".h", ".cuh")):
                    deps.append(os.path.join(dd, fn))
    for s in deps:
        with open(s, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("native build failed: " + " ".join(cmd[:3]))
    if os.environ.get("DEEPREC_BUILD_VERBOSE"):
        sys.stderr.write(r.stdout + r.stderr)


def _up_to_date(out: str, stamp: str) -> bool:
    sf = out + ".stamp"
    return os.path.exists(out) and os.path.exists(sf) and open(sf).read() == stamp


def build_host(force: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libdeeprec_host.so")
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    # -fopenmp + DR_USE_OPENMP: the host engine's parallel loops share PyTorch's libgomp pool (see ThreadPool in host_engine.cc)
    flags = ["-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-march=x86-64-v3", "-fno-math-errno", "-fopenmp", "-DDR_USE_OPENMP"]
    stamp = _stamp(srcs, flags)
    if not force and _up_to_date(out, stamp):
        return out
    _run(["g++"] + flags + srcs + ["-o", out, "-ldl"])
    open(out + ".stamp", "w").write(stamp)
    return out


def build_cuda(force: bool = False, verbose_ptxas: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    out = os.path.join(LIB, "libdeeprec_cuda.so")
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    flags = ARCH_FLAGS + ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC",
                          "--expt-relaxed-constexpr", "-I" + os.path.join(CSRC)]
    if verbose_ptxas:
        flags += ["-Xptxas", "-v"]
    stamp = _stamp(srcs, flags)
    if not force and _up_to_date(out, stamp):
        return out

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        ostamp = _stamp([src], flags)
        if not force and _up_to_date(obj, ostamp):
            return obj
        _run([NVCC] + flags + ["-c", src, "-o", obj])
        open(obj + ".stamp", "w").write(ostamp)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    _run([NVCC] + ARCH_FLAGS + ["-shared", "-Xcompiler", "-fPIC", "-cudart", "static"] + objs + ["-o", out])
    open(out + ".stamp", "w").write(stamp)
    return out


def build_all(force: bool = False) -> None:
    build_host(force)
    build_cuda(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", os.listdir(LIB))
