"""The C-ABI library loads on a CPU-only box and exports every symbol include/brotli_b200.h declares
(no compute calls without a GPU)."""
import os
import re

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from brotli_rs_b200 import _lib, build
    if not os.path.exists(_lib.library_path()):
        build.build()
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "brotli_b200.h")).read()
    declared = set(re.findall(r"\b(bro_[a-z_]+)\s*\(", header)) - {"bro_read_cb"}
    assert declared == set(_lib.ABI_SYMBOLS), declared ^ set(_lib.ABI_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_status_descriptions_match_reference_strings():
    """src/lib.rs:331-354, byte-identical (the oracle carries the same table; both restate the reference)."""
    from brotli_rs_b200 import status_description
    from oracle import oracle
    for st in list(range(0, 25)):
        assert status_description(st) == oracle.description(st)
    assert status_description(15) == "Enocuntered non-zero bit trailing the stream"


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle or the host simulation."""
    pkg = os.path.join(ROOT, "brotli_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".c")) and f != "bro_hostsim.cpp":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
                if f.endswith((".cu",)):
                    assert "BRO_HOSTSIM 1" not in text


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from brotli_rs_b200 import BatchDecoder, Decompressor
    with pytest.raises(RuntimeError):
        BatchDecoder()
    with pytest.raises(RuntimeError):
        Decompressor(b"\x06")


def test_sharding_balances_bytes():
    import numpy as np
    from brotli_rs_b200 import shard_streams
    rng = np.random.default_rng(0)
    in_lens = rng.integers(1, 400000, 5200)
    out_lens = in_lens * rng.integers(1, 60, 5200)
    for ws in (1, 2, 4, 8):
        shards = shard_streams(in_lens, out_lens, ws)
        allidx = np.sort(np.concatenate(shards))
        assert (allidx == np.arange(5200)).all()
        tot = np.array([out_lens[s].sum() + in_lens[s].sum() for s in shards], dtype=np.float64)
        assert tot.max() / tot.mean() < 1.02


def test_mg_partition_balances_contiguous_ranges():
    """bro_mg_partition (the split bro_mg_decode_host uses): contiguous, complete, balanced in compressed + slot bytes;
    host arithmetic only, so it runs without a GPU."""
    import numpy as np
    from brotli_rs_b200 import mg_partition
    rng = np.random.default_rng(3)
    for n in (1, 7, 5000):
        in_lens = rng.integers(1, 5000, n)
        caps = in_lens * rng.integers(1, 70, n)
        in_off = np.concatenate([[11], 11 + np.cumsum(in_lens)]).astype(np.uint64)       # offsets need not start at 0
        out_off = np.concatenate([[5], 5 + np.cumsum(caps)]).astype(np.uint64)
        for g in (1, 2, 3, 8):
            first = mg_partition(in_off, out_off, g)
            assert first[0] == 0 and first[-1] == n and (np.diff(first.astype(np.int64)) >= 0).all()
            if n >= 1000:
                work = np.array([float(in_lens[first[k]:first[k + 1]].sum() + caps[first[k]:first[k + 1]].sum()) for k in range(g)])
                assert work.max() / work.mean() < 1.05


def test_bro_resume_layout_matches_the_shim(tmp_path):
    """bro_resume crosses the ABI by value (arrays of it in device memory): its C layout must be what the Rust shim in
    INTEGRATION.md and BatchDecoder.RESUME_DTYPE declare -- 48 bytes, the offsets below."""
    import subprocess
    import numpy as np
    from brotli_rs_b200 import BatchDecoder
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "brotli_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(bro_resume), offsetof(bro_resume, in_bits), '
                   'offsetof(bro_resume, pos), offsetof(bro_resume, window), offsetof(bro_resume, dist), offsetof(bro_resume, p1), '
                   'offsetof(bro_resume, p2), offsetof(bro_resume, flags), offsetof(bro_resume, reserved)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [48, 0, 8, 12, 16, 32, 36, 40, 44]
    dt = BatchDecoder.RESUME_DTYPE
    assert dt.itemsize == 48 and [dt.fields[k][1] for k in ("in_bits", "pos", "window", "dist", "p1", "p2", "flags", "reserved")] == got[1:]
    shim = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert "pub struct bro_resume" in shim and "pub in_bits: u64" in shim and "pub dist: [u32; 4]" in shim
