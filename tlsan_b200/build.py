"""In-tree nvcc build of the C-ABI library (sm_100a only; cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("TLSAN_LIB") or os.path.join(HERE, "libtlsan_b200.so")    # TLSAN_LIB: A/B builds of the same sources
SOURCES = ["tlsan_abi.cu", "tlsan_fwd_bwd.cu", "tlsan_fused_mma.cu", "tlsan_fused_async.cu", "tlsan_fused_pf.cu", "tlsan_sort.cu", "tlsan_update.cu", "tlsan_dataset.cu", "tlsan_shard.cu", "tlsan_builder.cu", "tlsan_rank_tc.cu", "tlsan_host.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "128"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "tlsan_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> tlsan_b200/libtlsan_b200.so.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("TLSAN_NVCC_EXTRA", "").split()          # e.g. -DTLSAN_DEBUG_DWD for kernel debugging
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
