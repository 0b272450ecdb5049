"""Build libcpc_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C-ABI .so)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcpc_b200.so")
SOURCES = ["api.cu", "gemm_simt.cu", "gemm_tc.cu", "encoder.cu", "gru.cu", "criterion.cu", "score_mma.cu", "gru_mma.cu", "gru_mma_wide.cu", "lstm_mma.cu", "conv0_mma.cu", "thead.cu", "feeder.cu", "attn_mma.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--use_fast_math=false", "-Xptxas", "-v"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]
if os.environ.get("CPC_B200_TIMELINE") == "1":  # debug build: per-CTA cycle stamps in the persistent GEMM (tools/gemm_probe.py)
    FLAGS.append("-DCPC_B200_TIMELINE")


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "cpc_b200.h"), __file__]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, src):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        return job, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        for (src, obj), r in ex.map(compile_one, jobs):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                sys.stderr.write(r.stderr)
            else:
                for line in r.stderr.splitlines():
                    if "warning" in line or "spill" in line and "0 bytes spill" not in line:
                        sys.stderr.write(line + "\n")
    if jobs or not os.path.exists(OUT):
        r = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
