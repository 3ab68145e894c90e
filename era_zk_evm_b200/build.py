"""Builds the CUDA extension in-tree: era_zk_evm_b200/libzkb.so (sm_100a only, no fallback)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libzkb.so")
SOURCES = ["zkb.cu", "host_codec.cpp"]
DEPS = ["zkb.cu", "host_codec.cpp", "codec.cuh", "logsort.cuh", "comm.cuh", "consume.cuh", "alubench.cuh", "vm.cuh", "u256.cuh", "keccak.cuh", "sha256.cuh", "secp256k1.cuh", "isa_tables.inc"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared", "-ldl"]


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.join(HERE, "..", "include", h) for h in ("zkb.h", "zkb_records.h", "zkb_codec.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra=(), out: str = OUT) -> str:
    from . import isa
    isa.write_header()
    if force or out != OUT or stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + \
            ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.check_call(cmd)
    return out


def build_variant(tag: str, warps: int, min_ctas: int, extra=()) -> str:
    """kernel-tuning experiments: build/variants/libzkb_<tag>.so (selected with ZKB_LIB_PATH)"""
    d = os.path.join(HERE, "..", "build", "variants")
    os.makedirs(d, exist_ok=True)
    out = os.path.abspath(os.path.join(d, f"libzkb_{tag}.so"))
    return build(force=True, extra=[f"-DZKB_WARPS_PER_CTA={warps}", f"-DZKB_MIN_CTAS_PER_SM={min_ctas}"] + list(extra), out=out)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
