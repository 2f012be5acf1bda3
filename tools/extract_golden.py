#!/usr/bin/env python3
"""Build tests/golden/ from the reference's own test material (run in the authoring container).

/root/reference does not exist on the GPU box, so everything the parity tests need travels as
fixtures committed under tests/golden/:

  tests/golden/data/            the reference's data/ corpus (binary fixtures, Apache-2.0), verbatim
  tests/golden/stream_vectors.json   the 33 stream tests of tests/lib.rs:4-605 (+ doc-test src/lib.rs:361-376):
                                input bytes (hex) or data file, expected output (text, data file) or
                                the expected error substring of #[should_panic(expected=...)]
  tests/golden/transform_kats.json   the 121 transform KATs of src/transformation/mod.rs:215-1301
  tests/golden/bitreader_kats.json   bit-reader KATs of src/bitreader/mod.rs:339-560 (as op scripts)
  tests/golden/imtf_kats.json        the IMTF vectors of tests/lib.rs:653-673

Only *test vectors* (inputs and expected outputs) are extracted; no reference code is copied.
"""
import json
import os
import re
import shutil
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rust_unescape(s):
    """Decode a Rust (byte) string literal body to bytes."""
    out = bytearray()
    i = 0
    while i < len(s):
        c = s[i]
        if c == "\\":
            n = s[i + 1]
            if n == "x":
                out.append(int(s[i + 2:i + 4], 16))
                i += 4
                continue
            if n == "u":
                j = s.index("}", i)
                out += chr(int(s[i + 3:j], 16)).encode("utf-8")
                i = j + 1
                continue
            out += {"n": b"\n", "t": b"\t", "r": b"\r", "\\": b"\\", '"': b'"', "'": b"'", "0": b"\0"}[n]
            i += 2
            continue
        out += c.encode("utf-8")
        i += 1
    return bytes(out)


def split_tests(text):
    """Yield (attrs, name, body) for every #[test] fn."""
    for m in re.finditer(r"#\[test\](.*?)fn\s+(\w+)\s*\(\)\s*\{(.*?)\n\}", text, re.S):
        yield m.group(1), m.group(2), m.group(3)


def stream_vectors():
    text = open(os.path.join(REF, "tests/lib.rs"), encoding="utf-8").read()
    vecs = []
    for attrs, name, body in split_tests(text):
        if "Decompressor" not in body:
            continue
        v = {"name": name, "source": "tests/lib.rs"}
        m = re.search(r'should_panic\(expected\s*=\s*"([^"]*)"\)', attrs)
        if m:
            v["expect_error_substring"] = m.group(1)
        # input
        m = re.search(r"Cursor::new\(vec!\[(.*?)\]\)", body, re.S)
        if m:
            v["input_hex"] = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1))).hex()
        else:
            m = re.search(r'&b"((?:[^"\\]|\\.)*)"\.to_vec\(\)', body, re.S)
            if m:
                v["input_hex"] = rust_unescape(m.group(1)).hex()
            else:
                m = re.search(r'brotli_stream\s*=\s*(?:std::fs::)?File::open\("data/([^"]+)"\)', body)
                assert m, name
                v["input_file"] = m.group(1)
        # expectation
        if "expect_error_substring" not in v:
            m = re.search(r'assert_eq!\("((?:[^"\\]|\\.)*)",\s*decompressed\)', body, re.S)
            if m:
                v["expect_hex"] = rust_unescape(m.group(1)).hex()
            else:
                m = re.search(r'File::open\("data/([^"]+)"\)\.unwrap\(\)\.read_to_(?:string|end)\(&mut expected\)', body)
                assert m, name
                v["expect_file"] = m.group(1)
        vecs.append(v)
    # doc-test src/lib.rs:361-376
    vecs.append({"name": "doctest_64x", "source": "src/lib.rs:361-376", "input_file": "64x.compressed", "expect_file": "64x"})
    return vecs


def transform_kats():
    text = open(os.path.join(REF, "src/transformation/mod.rs"), encoding="utf-8").read()
    text = text[text.index("mod tests"):]
    kats = []
    for m in re.finditer(r"fn\s+(\w+)\s*\(\)\s*\{(.*?)\n\t\}", text, re.S):
        name, body = m.group(1), m.group(2)
        mt = re.search(r"transformation\((\d+),", body)
        if not mt:
            continue
        tid = int(mt.group(1))
        mb = re.search(r'let\s+base_word\s*=\s*String::from\("((?:[^"\\]|\\.)*)"\)', body, re.S)
        assert mb, (name, body)
        base = rust_unescape(mb.group(1))
        me = re.search(r'let\s+expected\s*=\s*"((?:[^"\\]|\\.)*)"', body, re.S)
        if me:
            exp = rust_unescape(me.group(1))
        else:
            # test 102 (src/transformation/mod.rs:1133-1139): expected = [vec![bytes], base_word].concat()
            me = re.search(r"let\s+expected\s*=\s*\[vec!\[(.*?)\],\s*base_word\.clone\(\)\]\.concat\(\)", body, re.S)
            assert me, (name, body)
            exp = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", me.group(1))) + base
        kats.append({"name": name, "id": tid, "base_hex": base.hex(), "expect_hex": exp.hex()})
    return kats


def imtf_kats():
    text = open(os.path.join(REF, "tests/lib.rs"), encoding="utf-8").read()
    out = []
    for attrs, name, body in split_tests(text):
        if name not in ("should_not_change", "should_compose_to_identity"):
            continue
        m = re.search(r"vec!\[(.*?)\]", body, re.S)
        v = [int(x) for x in re.findall(r"\d+", m.group(1))]
        out.append({"name": name, "vector": v,
                    "property": "imtf(v) == v" if name == "should_not_change" else "mtf(imtf(v)) == v"})
    return out


def bitreader_kats():
    """The 13 bit-reader unit tests of src/bitreader/mod.rs:339-560, restated by hand as op scripts
    (op, argument, expected value or None).  Ops: u8, bit, bits n, tail, nibble, nibbles n, string n."""
    x5, y5 = [0x78] * 5, [0x79] * 5
    fname = x5 + y5 + [0x2e, 0x74, 0x78, 0x74, 0x00]
    kats = [
        ("should_read_one_u8", [0x1f, 0x8b], [["u8", None, 0x1f]]),
        ("should_read_two_u8", [0x1f, 0x8b], [["u8", None, 0x1f], ["u8", None, 0x8b]]),
        ("should_read_one_set_bit", [3], [["bit", None, 1]]),
        ("should_read_some_bits", [134, 1], [["bit", None, b] for b in (0, 1, 1, 0, 0, 0, 0, 1, 1, 0)]),
        ("should_read_u8_after_bit", [0b10001101, 0b00010101], [["bit", None, 1], ["u8", None, 0b11000110]]),
        ("should_read_fixed_length_string", fname + fname, [["string", 14, b"xxxxxyyyyy.txt".hex()]]),
        ("should_read_29u8_from_5_bits", [157], [["bits", 5, 29]]),
        ("should_read_3784u16_from_11_bits", [0b11001000, 0b11111110], [["bits", 11, 1736]]),
        ("should_read_19u8_from_byte_tail", [0b10011101], [["bit", None, None]] * 3 + [["tail", None, 19]]),
        ("should_read_10u8_from_nibble", [0b11010101], [["bit", None, None]] * 3 + [["nibble", None, 10]]),
        ("should_read_10u8_nibble_twice", [0b10101010], [["nibble", None, 10], ["nibble", None, 10]]),
        ("should_read_7u8_nibble_four_times", [0b11101111, 0b11101110, 0b11101110],
         [["bit", None, None]] + [["nibble", None, 7]] * 4),
        ("should_read_524527u32_from_5_nibbles", [0b11101111, 0, 0b11101000], [["nibbles", 5, 524527]]),
    ]
    return [{"name": n, "data_hex": bytes(d).hex(), "ops": o} for n, d, o in kats]


def main():
    if not os.path.isdir(REF):
        print("reference not present; nothing to do")
        return 1
    os.makedirs(os.path.join(GOLD, "data"), exist_ok=True)
    n = 0
    for fn in sorted(os.listdir(os.path.join(REF, "data"))):
        if fn.startswith("."):
            continue
        shutil.copyfile(os.path.join(REF, "data", fn), os.path.join(GOLD, "data", fn))
        os.chmod(os.path.join(GOLD, "data", fn), 0o644)
        n += 1
    print("copied", n, "corpus files")
    sv = stream_vectors()
    json.dump(sv, open(os.path.join(GOLD, "stream_vectors.json"), "w"), indent=1)
    print("stream vectors:", len(sv), "(", sum("expect_error_substring" in v for v in sv), "must-reject )")
    tk = transform_kats()
    json.dump(tk, open(os.path.join(GOLD, "transform_kats.json"), "w"), indent=1)
    print("transform KATs:", len(tk), "ids", len(set(k["id"] for k in tk)))
    bk = bitreader_kats()
    json.dump(bk, open(os.path.join(GOLD, "bitreader_kats.json"), "w"), indent=1)
    print("bitreader KATs:", len(bk))
    ik = imtf_kats()
    json.dump(ik, open(os.path.join(GOLD, "imtf_kats.json"), "w"), indent=1)
    print("imtf KATs:", len(ik))
    return 0


if __name__ == "__main__":
    sys.exit(main())
