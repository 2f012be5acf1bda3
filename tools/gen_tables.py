#!/usr/bin/env python3
"""Generate the static Brotli format tables used by the CUDA decoder and by the oracle.

Everything produced here is *format data* defined by the Brotli specification
(draft-alakuijala-brotli-07 / RFC 7932), not code.  Each table is re-derived from its
rule in the specification and then checked against the check value the specification
itself publishes, so no table is transcribed from the reference's sources:

  * context lookup tables Lut0/Lut1/Lut2     spec section 7.1, CRC-32 values given there
      (reference: src/lookuptable/mod.rs:1-56)
  * insert/copy length code table             spec section 5
      (reference: src/lookuptable/mod.rs:61-123)
  * block count code table                    spec section 6
      (reference: src/lib.rs:962-976)
  * 121 word transforms                       spec appendix B, 648-byte image CRC 0x3d965f81
      (reference: src/transformation/mod.rs:84-209)
  * static dictionary (122,784 bytes)         spec appendix A, CRC-32 0x5136cb04
      (reference: src/dictionary/mod.rs:1-13); the bytes are taken from the system's
      libbrotlicommon (BrotliGetDictionary) and accepted only if the CRC matches.

Outputs:
  brotli_rs_b200/csrc/bro_tables_generated.h
  brotli_rs_b200/data/dictionary.bin

When /root/reference is present the script additionally cross-checks every table against
the reference's own source arrays (parsed as text; nothing is copied).
"""
import ctypes
import hashlib
import os
import re
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_H = os.path.join(ROOT, "brotli_rs_b200", "csrc", "bro_tables_generated.h")
OUT_DICT = os.path.join(ROOT, "brotli_rs_b200", "data", "dictionary.bin")
REF = "/root/reference"


# ----------------------------------------------------------------------------- context LUTs
def make_luts():
    vowels_up = set(b"AEIOU")
    vowels_lo = set(b"aeiou")
    lut0 = [0] * 256
    for c in (9, 10, 13):
        lut0[c] = 4
    row = [8, 12, 16, 12, 12, 20, 12, 16, 24, 28, 12, 12, 32, 12, 36, 12] + [44] * 10 + [32, 32, 24, 40, 28, 12]
    lut0[32:64] = row
    lut0[64] = 12
    for c in range(65, 91):
        lut0[c] = 48 if c in vowels_up else 52
    lut0[91:96] = [24, 12, 28, 12, 12]
    lut0[96] = 12
    for c in range(97, 123):
        lut0[c] = 56 if c in vowels_lo else 60
    lut0[123:127] = [24, 12, 28, 12]
    lut0[127] = 0
    for c in range(128, 192):
        lut0[c] = c & 1
    for c in range(192, 256):
        lut0[c] = 2 + (c & 1)

    lut1 = [0] * 256
    for c in range(33, 127):
        if 48 <= c <= 57 or 65 <= c <= 90:
            lut1[c] = 2
        elif 97 <= c <= 122:
            lut1[c] = 3
        else:
            lut1[c] = 1
    for c in range(224, 256):
        lut1[c] = 2

    lut2 = [0] * 256
    for c in range(256):
        if c == 0:
            v = 0
        elif c < 16:
            v = 1
        elif c < 64:
            v = 2
        elif c < 128:
            v = 3
        elif c < 192:
            v = 4
        elif c < 240:
            v = 5
        elif c < 255:
            v = 6
        else:
            v = 7
        lut2[c] = v

    assert zlib.crc32(bytes(lut0)) == 0x8E91EFB7, "Lut0 CRC mismatch vs spec"
    assert zlib.crc32(bytes(lut1)) == 0xD01A32F4, "Lut1 CRC mismatch vs spec"
    assert zlib.crc32(bytes(lut2)) == 0x0DD7A0D6, "Lut2 CRC mismatch vs spec"
    return lut0, lut1, lut2


# ----------------------------------------------------------------------------- insert / copy codes
INS_BASE_EXTRA = None
COPY_BASE_EXTRA = None


def make_insert_copy():
    ins_extra = [0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24]
    cpy_extra = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 24]
    ins, b = [], 0
    for e in ins_extra:
        ins.append((b, e))
        b += 1 << e
    cpy, b = [], 2
    for e in cpy_extra:
        cpy.append((b, e))
        b += 1 << e
    assert ins[23] == (22594, 24) and cpy[23] == (2118, 24) and ins[16] == (130, 6) and cpy[18] == (134, 6)
    cell_ins = [0, 0, 0, 0, 8, 8, 0, 16, 8, 16, 16]
    cell_cpy = [0, 8, 0, 8, 0, 8, 16, 0, 16, 8, 16]
    table = []
    for sym in range(704):
        cell = sym >> 6
        ic = cell_ins[cell] + ((sym >> 3) & 7)
        cc = cell_cpy[cell] + (sym & 7)
        table.append((ins[ic][0], ins[ic][1], cpy[cc][0], cpy[cc][1]))
    return table


def make_block_counts():
    extra = [2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12, 13, 24]
    out, b = [], 1
    for e in extra:
        out.append((b, e))
        b += 1 << e
    assert out[25] == (16625, 24) and out[18] == (369, 7) and out[4] == (17, 3)
    return out


# ----------------------------------------------------------------------------- transforms
ID, UF, UA = 0, 1, 2          # Identity, UppercaseFirst, UppercaseAll
OF = lambda n: 2 + n          # OmitFirst1..9 -> 3..11
OL = lambda n: 11 + n         # OmitLast1..9  -> 12..20

TRANSFORMS = [
    (b"", ID, b""), (b"", ID, b" "), (b" ", ID, b" "), (b"", OF(1), b""), (b"", UF, b" "),
    (b"", ID, b" the "), (b" ", ID, b""), (b"s ", ID, b" "), (b"", ID, b" of "), (b"", UF, b""),
    (b"", ID, b" and "), (b"", OF(2), b""), (b"", OL(1), b""), (b", ", ID, b" "), (b"", ID, b", "),
    (b" ", UF, b" "), (b"", ID, b" in "), (b"", ID, b" to "), (b"e ", ID, b" "), (b"", ID, b"\""),
    (b"", ID, b"."), (b"", ID, b"\">"), (b"", ID, b"\n"), (b"", OL(3), b""), (b"", ID, b"]"),
    (b"", ID, b" for "), (b"", OF(3), b""), (b"", OL(2), b""), (b"", ID, b" a "), (b"", ID, b" that "),
    (b" ", UF, b""), (b"", ID, b". "), (b".", ID, b""), (b" ", ID, b", "), (b"", OF(4), b""),
    (b"", ID, b" with "), (b"", ID, b"'"), (b"", ID, b" from "), (b"", ID, b" by "), (b"", OF(5), b""),
    (b"", OF(6), b""), (b" the ", ID, b""), (b"", OL(4), b""), (b"", ID, b". The "), (b"", UA, b""),
    (b"", ID, b" on "), (b"", ID, b" as "), (b"", ID, b" is "), (b"", OL(7), b""), (b"", OL(1), b"ing "),
    (b"", ID, b"\n\t"), (b"", ID, b":"), (b" ", ID, b". "), (b"", ID, b"ed "), (b"", OF(9), b""),
    (b"", OF(7), b""), (b"", OL(6), b""), (b"", ID, b"("), (b"", UF, b", "), (b"", OL(8), b""),
    (b"", ID, b" at "), (b"", ID, b"ly "), (b" the ", ID, b" of "), (b"", OL(5), b""), (b"", OL(9), b""),
    (b" ", UF, b", "), (b"", UF, b"\""), (b".", ID, b"("), (b"", UA, b" "), (b"", UF, b"\">"),
    (b"", ID, b"=\""), (b" ", ID, b"."), (b".com/", ID, b""), (b" the ", ID, b" of the "), (b"", UF, b"'"),
    (b"", ID, b". This "), (b"", ID, b","), (b".", ID, b" "), (b"", UF, b"("), (b"", UF, b"."),
    (b"", ID, b" not "), (b" ", ID, b"=\""), (b"", ID, b"er "), (b" ", UA, b" "), (b"", ID, b"al "),
    (b" ", UA, b""), (b"", ID, b"='"), (b"", UA, b"\""), (b"", UF, b". "), (b" ", ID, b"("),
    (b"", ID, b"ful "), (b" ", UF, b". "), (b"", ID, b"ive "), (b"", ID, b"less "), (b"", UA, b"'"),
    (b"", ID, b"est "), (b" ", UF, b"."), (b"", UA, b"\">"), (b" ", ID, b"='"), (b"", UF, b","),
    (b"", ID, b"ize "), (b"", UA, b"."), (b"\xc2\xa0", ID, b""), (b" ", ID, b","), (b"", UF, b"=\""),
    (b"", UA, b"=\""), (b"", ID, b"ous "), (b"", UA, b", "), (b"", UF, b"='"), (b" ", UF, b","),
    (b" ", UA, b"=\""), (b" ", UA, b", "), (b"", UA, b","), (b"", UA, b"("), (b"", UA, b". "),
    (b" ", UA, b"."), (b"", UA, b"='"), (b" ", UA, b". "), (b" ", UF, b"=\""), (b" ", UA, b"='"),
    (b" ", UF, b"='"),
]


def check_transforms():
    assert len(TRANSFORMS) == 121
    img = b"".join(p + b"\0" + bytes([t]) + s + b"\0" for p, t, s in TRANSFORMS)
    assert len(img) == 648, len(img)
    assert zlib.crc32(img) == 0x3D965F81, "transform image CRC mismatch vs spec appendix B"


# ----------------------------------------------------------------------------- dictionary
class _BrotliDictionary(ctypes.Structure):
    _fields_ = [
        ("size_bits_by_length", ctypes.c_uint8 * 32),
        ("offsets_by_length", ctypes.c_uint32 * 32),
        ("data_size", ctypes.c_size_t),
        ("data", ctypes.POINTER(ctypes.c_uint8)),
    ]


def load_dictionary():
    lib = ctypes.CDLL("libbrotlicommon.so.1")
    lib.BrotliGetDictionary.restype = ctypes.POINTER(_BrotliDictionary)
    d = lib.BrotliGetDictionary().contents
    assert d.data_size == 122784, d.data_size
    data = bytes(ctypes.string_at(d.data, d.data_size))
    assert zlib.crc32(data) == 0x5136CB04, "dictionary CRC mismatch vs spec appendix A"
    bits = list(d.size_bits_by_length)[:25]
    offs = list(d.offsets_by_length)[:25]
    # spec section 8: offsets are the running sum of (length << size_bits) over lengths 4..24
    run, chk = 0, [0] * 25
    for length in range(25):
        chk[length] = run
        if bits[length]:
            run += length << bits[length]
    assert run == 122784 and chk == offs, (chk, offs)
    return data, bits, offs


# ----------------------------------------------------------------------------- reference cross-check
def _rust_array(text, name):
    m = re.search(name + r"[^=]*=\s*\[(.*?)\];", text, re.S)
    assert m, name
    return m.group(1)


def crosscheck_reference(lut0, lut1, lut2, ic, data, bits, offs):
    if not os.path.isdir(REF):
        print("reference not present: cross-check skipped")
        return
    lt = open(os.path.join(REF, "src/lookuptable/mod.rs")).read()
    for name, mine in (("LUT_0", lut0), ("LUT_1", lut1), ("LUT_2", lut2)):
        ref = [int(x) for x in re.findall(r"\d+", _rust_array(lt, "pub const " + name))]
        # first number matched is the '256' of the type annotation when present
        ref = ref[-256:]
        assert ref == mine, name
    nums = [int(x) for x in re.findall(r"\d+", _rust_array(lt, "pub const INSERT_LENGTHS_AND_COPY_LENGTHS"))]
    nums = nums[-704 * 4:]
    flat = [v for row in ic for v in row]
    assert nums == flat, "insert/copy table differs from reference"
    dt = open(os.path.join(REF, "src/dictionary/mod.rs")).read()
    ref_offs = [int(x) for x in re.findall(r"\d+", _rust_array(dt, "pub const BROTLI_DICTIONARY_OFFSETS_BY_LENGTH"))][-25:]
    ref_bits = [int(x) for x in re.findall(r"\d+", _rust_array(dt, "pub const BROTLI_DICTIONARY_SIZE_BITS_BY_LENGTH"))][-25:]
    assert ref_offs == offs and ref_bits == bits
    body = _rust_array(dt, r"pub const BROTLI_DICTIONARY:")
    ref_data = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", body))
    assert ref_data == data, "dictionary differs from reference"
    print("reference cross-check OK (LUTs, insert/copy table, dictionary %s)" % hashlib.sha256(data).hexdigest()[:16])


# ----------------------------------------------------------------------------- emit
def c_array(name, ctype, vals, per_line=16):
    lines = []
    for i in range(0, len(vals), per_line):
        lines.append("  " + ", ".join(str(v) for v in vals[i:i + per_line]) + ",")
    return "BRO_TABLE_QUAL %s %s[%d] = {\n%s\n};\n" % (ctype, name, len(vals), "\n".join(lines))


def main():
    lut0, lut1, lut2 = make_luts()
    ic = make_insert_copy()
    bc = make_block_counts()
    check_transforms()
    data, bits, offs = load_dictionary()
    crosscheck_reference(lut0, lut1, lut2, ic, data, bits, offs)

    # transforms -> packed strings + per-transform descriptor
    blob = bytearray()
    desc = []
    for p, t, s in TRANSFORMS:
        po = len(blob)
        blob += p
        so = len(blob)
        blob += s
        desc.append((po, len(p), t, so, len(s)))
    assert len(blob) < 65536

    out = []
    out.append("// GENERATED by tools/gen_tables.py -- do not edit.\n")
    out.append("// Format constants of the Brotli specification (draft-alakuijala-brotli-07), each checked\n"
               "// against the check value the specification publishes.  See the generator for provenance.\n")
    out.append("#pragma once\n#include <stdint.h>\n")
    out.append("#ifndef BRO_TABLE_QUAL\n#define BRO_TABLE_QUAL static const\n#endif\n\n")
    out.append("// spec 7.1 (reference src/lookuptable/mod.rs:1-56)\n")
    out.append(c_array("bro_lut0", "uint8_t", lut0))
    out.append(c_array("bro_lut1", "uint8_t", lut1))
    out.append(c_array("bro_lut2", "uint8_t", lut2))
    out.append("// spec 5 (reference src/lookuptable/mod.rs:123): per insert&copy symbol\n"
               "// packed as ins_base | ins_extra<<16 and copy_base | copy_extra<<16\n")
    out.append(c_array("bro_ic_insert", "uint32_t", [r[0] | (r[1] << 16) for r in ic], 8))
    out.append(c_array("bro_ic_copy", "uint32_t", [r[2] | (r[3] << 16) for r in ic], 8))
    out.append("// spec 6 (reference src/lib.rs:962-976): block count code base | extra<<16\n")
    out.append(c_array("bro_block_count", "uint32_t", [b | (e << 16) for b, e in bc], 8))
    out.append("// spec 8 / appendix A (reference src/dictionary/mod.rs:1-11)\n")
    out.append(c_array("bro_dict_offsets", "uint32_t", offs))
    out.append(c_array("bro_dict_size_bits", "uint8_t", bits))
    out.append("#define BRO_DICT_SIZE 122784\n\n")
    out.append("// spec appendix B (reference src/transformation/mod.rs:84-209)\n"
               "// type: 0 identity, 1 uppercase-first, 2 uppercase-all, 3..11 omit-first 1..9, 12..20 omit-last 1..9\n")
    out.append(c_array("bro_xf_strings", "uint8_t", list(blob), 24))
    out.append(c_array("bro_xf_prefix_off", "uint16_t", [d[0] for d in desc]))
    out.append(c_array("bro_xf_prefix_len", "uint8_t", [d[1] for d in desc]))
    out.append(c_array("bro_xf_type", "uint8_t", [d[2] for d in desc]))
    out.append(c_array("bro_xf_suffix_off", "uint16_t", [d[3] for d in desc]))
    out.append(c_array("bro_xf_suffix_len", "uint8_t", [d[4] for d in desc]))

    os.makedirs(os.path.dirname(OUT_H), exist_ok=True)
    os.makedirs(os.path.dirname(OUT_DICT), exist_ok=True)
    with open(OUT_H, "w") as f:
        f.write("".join(out))
    with open(OUT_DICT, "wb") as f:
        f.write(data)
    print("wrote", OUT_H)
    print("wrote", OUT_DICT, "sha256", hashlib.sha256(data).hexdigest())


if __name__ == "__main__":
    sys.exit(main())
