"""Builds libbrotli_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m brotli_rs_b200.build

The library is pure CUDA runtime + C ABI (include/brotli_b200.h); it does not link against torch.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libbrotli_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
SOURCES = ["bro_kernels.cu", "bro_kernels_parse.cu", "bro_kernels_copy.cu", "bro_kernels_resume.cu", "bro_abi.cu"]
BLOB = "dict_blob.c"
HEADERS = ["bro_decoder_core.h", "bro_parse.h", "bro_records.h", "bro_copy_piece.h", "bro_kernels.h", "bro_status.h", "bro_tables_generated.h"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA decoder cannot be built (there is no CPU fallback)")
    return exe


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS + [BLOB]] + [os.path.join(HERE, "data", "dictionary.bin"),
                                                                   os.path.join(HERE, "..", "include", "brotli_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an experimental variant (e.g. defines=["BRO_MIN_BLOCKS=3"], out="libvariant.so")."""
    global LIB
    if out is not None:
        LIB_SAVED, LIB = LIB, os.path.join(LIBDIR, out)
        try:
            return _build(True, verbose, defines)
        finally:
            LIB = LIB_SAVED
    return _build(force, verbose, defines)


def _build(force, verbose, defines):
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    # the dictionary image is embedded with .incbin by the host assembler (nvcc would split -Wa,... at the comma)
    blob_o = os.path.join(LIBDIR, "dict_blob.o")
    subprocess.check_call([os.environ.get("CC", "gcc"), "-c", "-fPIC", os.path.join(CSRC, BLOB),
                           "-Wa,-I" + os.path.join(HERE, "data"), "-o", blob_o])
    cmd = [nvcc()] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                             "-Xptxas", "-v", "-o", LIB] + ["-D" + d for d in defines] + \
          [os.path.join(CSRC, s) for s in SOURCES] + [blob_o]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libbrotli_b200.so")
    with open(os.path.join(LIBDIR, os.path.basename(LIB) + ".ptxas_info.txt"), "w") as f:
        f.write("".join(l for l in res.stderr.splitlines(True) if "Compile time" not in l))   # stable across builds
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
