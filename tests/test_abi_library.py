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


def _parse_c_array(text, name):
    import re
    m = re.search(r"\b%s\[\d+\]\s*=\s*\{(.*?)\};" % name, text, re.S)
    assert m, name
    return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]


def test_committed_tables_match_the_specification_check_values():
    """bro_tables_generated.h is included by the product AND by the oracle, so a wrong table would be wrong on both sides and
    the parity tests could not see it.  Each table of the committed header is therefore checked here against the check value
    the Brotli specification itself publishes (RFC 7932 section 7.1: CRC-32 of the three context look-up tables; appendix
    B: CRC-32 of the 648-byte transform image; appendix A: CRC-32 of the dictionary), and the insert / copy / block count
    code tables against the rules of sections 5 and 6 -- independently of tools/gen_tables.py, which made them."""
    import zlib
    text = open(os.path.join(ROOT, "brotli_rs_b200", "csrc", "bro_tables_generated.h")).read()
    for name, crc in (("bro_lut0", 0x8E91EFB7), ("bro_lut1", 0xD01A32F4), ("bro_lut2", 0x0DD7A0D6)):
        assert zlib.crc32(bytes(_parse_c_array(text, name))) == crc, name
    # transforms: rebuild the specification's image (prefix \0 type suffix \0 per transform) from the committed descriptors
    strings = bytes(_parse_c_array(text, "bro_xf_strings"))
    po, pl = _parse_c_array(text, "bro_xf_prefix_off"), _parse_c_array(text, "bro_xf_prefix_len")
    so, sl = _parse_c_array(text, "bro_xf_suffix_off"), _parse_c_array(text, "bro_xf_suffix_len")
    ty = _parse_c_array(text, "bro_xf_type")
    # transform types as the specification's image numbers them: 0 identity, 1 uppercase-first, 2 uppercase-all, 3..11 omit-first
    # 1..9, 12..20 omit-last 1..9 (the numbering bro_xf_type uses)
    img = b"".join(strings[po[i]: po[i] + pl[i]] + b"\0" + bytes([ty[i]]) + strings[so[i]: so[i] + sl[i]] + b"\0" for i in range(121))
    assert len(img) == 648 and zlib.crc32(img) == 0x3D965F81
    dic = open(os.path.join(ROOT, "brotli_rs_b200", "data", "dictionary.bin"), "rb").read()
    assert len(dic) == 122784 and zlib.crc32(dic) == 0x5136CB04
    bits = _parse_c_array(text, "bro_dict_size_bits")
    offs = _parse_c_array(text, "bro_dict_offsets")
    assert bits == [0, 0, 0, 0, 10, 10, 11, 11, 10, 10, 10, 10, 10, 9, 9, 8, 7, 7, 8, 7, 7, 6, 6, 5, 5]
    run = 0
    for length in range(25):
        assert offs[length] == run, length
        run += (length << bits[length]) if length >= 4 else 0
    assert run == 122784
    # section 5: insert and copy length codes (base, extra bits) and the cell layout of the 704 insert&copy symbols
    ins = [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 1), (8, 1), (10, 2), (14, 2), (18, 3), (26, 3), (34, 4), (50, 4), (66, 5), (98, 5),
           (130, 6), (194, 7), (322, 8), (578, 9), (1090, 10), (2114, 12), (6210, 14), (22594, 24)]
    cop = [(2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (7, 0), (8, 0), (9, 0), (10, 1), (12, 1), (14, 2), (18, 2), (22, 3), (30, 3), (38, 4), (54, 4),
           (70, 5), (102, 5), (134, 6), (198, 7), (326, 8), (582, 9), (1094, 10), (2118, 24)]
    for k in range(1, 24):      # a code's range ends where the next begins
        assert ins[k][0] == ins[k - 1][0] + (1 << ins[k - 1][1]) and cop[k][0] == cop[k - 1][0] + (1 << cop[k - 1][1])
    cells = [(0, 0), (0, 8), (0, 0), (0, 8), (8, 0), (8, 8), (0, 16), (16, 0), (8, 16), (16, 8), (16, 16)]
    ti, tc = _parse_c_array(text, "bro_ic_insert"), _parse_c_array(text, "bro_ic_copy")
    for sym in range(704):
        i0, c0 = cells[sym >> 6]
        ib, ie = ins[i0 + ((sym >> 3) & 7)]
        cb, ce = cop[c0 + (sym & 7)]
        assert ti[sym] == ib | (ie << 16) and tc[sym] == cb | (ce << 16), sym
    # section 6: block count codes
    bc = _parse_c_array(text, "bro_block_count")
    base = 1
    extra = [2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12, 13, 24]
    for k in range(26):
        assert bc[k] == base | (extra[k] << 16), k
        base += 1 << extra[k]
