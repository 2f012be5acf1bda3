"""The C++ Decompressor twin (include/brotli_b200.hpp) compiles against the C ABI on a CPU box, and on the GPU box
decodes the reference's doc-test (src/lib.rs:361-376) and rejects an invalid stream with the reference's message."""
import os
import subprocess

import pytest

from conftest import DATA, ROOT

SRC = r'''
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>
#include "brotli_b200.hpp"
static std::vector<uint8_t> slurp(const char* path) {
    std::ifstream g(path, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(g)), std::istreambuf_iterator<char>());
}
int main(int argc, char** argv) {
    if (argc > 3 && !std::strcmp(argv[1], "batch")) {
        // batch mode: argv[2..] = compressed files; prints "status size" per stream (no size hints given)
        bro_ctx* ctx = nullptr;
        if (bro_ctx_create(&ctx, -1) != BRO_OK) return 4;
        std::vector<std::vector<uint8_t>> streams;
        for (int i = 2; i < argc; i++) streams.push_back(slurp(argv[i]));
        for (const auto& it : brotli::decode_batch(ctx, streams)) std::printf("%d %zu\n", it.status, it.bytes.size());
        bro_ctx_destroy(ctx);
        return 0;
    }
    if (argc > 3 && !std::strcmp(argv[1], "mg")) {
        // multi-GPU batch mode: argv[2] = devices (0 = all), argv[3..] = compressed files, each decoded 40 times; slots sized by a
        // first unsized decode; prints the device count, then "status size" per distinct stream if every replica agrees
        bro_ctx* ctx = nullptr;
        if (bro_ctx_create(&ctx, -1) != BRO_OK) return 4;
        std::vector<std::vector<uint8_t>> streams;
        for (int i = 3; i < argc; i++) streams.push_back(slurp(argv[i]));
        const auto first = brotli::decode_batch(ctx, streams);
        bro_ctx_destroy(ctx);
        std::vector<std::vector<uint8_t>> many;
        std::vector<size_t> caps;
        for (int rep = 0; rep < 40; rep++)
            for (size_t i = 0; i < streams.size(); i++) { many.push_back(streams[i]); caps.push_back(first[i].bytes.size() + (first[i].status ? 4096 : 0)); }
        const auto res = brotli::decode_batch_multi_gpu(many, caps, std::atoi(argv[2]));
        for (size_t k = 0; k < res.size(); k++) {
            const auto& want = first[k % streams.size()];
            if (res[k].status != want.status || (want.status == 0 && res[k].bytes != want.bytes)) { std::printf("MISMATCH %zu\n", k); return 2; }
        }
        for (const auto& it : first) std::printf("%d %zu\n", it.status, it.bytes.size());
        return 0;
    }
    if (argc > 4 && !std::strcmp(argv[1], "stream")) {
        // streaming mode: argv[2] = compressed file, argv[3] = expected file, argv[4] = bytes asked of the file at a time
        std::ifstream f(argv[2], std::ios::binary);
        brotli::Decompressor<brotli::IstreamReader> d{brotli::IstreamReader(f), {static_cast<size_t>(std::atol(argv[4]))}};
        std::vector<uint8_t> out;
        try { d.read_to_end(out); } catch (const brotli::Error& e) { std::printf("ERR %d %s after %zu\n", e.status(), e.what(), out.size()); return 3; }
        const std::vector<uint8_t> exp = slurp(argv[3]);
        std::printf("%s %zu\n", out == exp ? "EQUAL" : "DIFFERENT", out.size());
        return out == exp ? 0 : 2;
    }
    std::ifstream f(argv[1], std::ios::binary);
    brotli::Decompressor<brotli::IstreamReader> d{brotli::IstreamReader(f)};
    std::vector<uint8_t> out;
    try { d.read_to_end(out); } catch (const brotli::Error& e) { std::printf("ERR %d %s\n", e.status(), e.what()); return 3; }
    std::ifstream g(argv[2], std::ios::binary);
    std::vector<uint8_t> exp((std::istreambuf_iterator<char>(g)), std::istreambuf_iterator<char>());
    std::printf("%s %zu\n", out == exp ? "EQUAL" : "DIFFERENT", out.size());
    return out == exp ? 0 : 2;
}
'''


def _build(tmp):
    from brotli_rs_b200 import _lib, build
    if not os.path.exists(_lib.library_path()):
        build.build()
    src = os.path.join(tmp, "twin.cpp")
    exe = os.path.join(tmp, "twin")
    open(src, "w").write(SRC)
    libdir = os.path.dirname(_lib.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L" + libdir, "-lbrotli_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_twin_compiles_and_links(tmp_path):
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
def test_cpp_twin_doctest_and_error(tmp_path):
    exe = _build(str(tmp_path))
    r = subprocess.run([exe, os.path.join(DATA, "64x.compressed"), os.path.join(DATA, "64x")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("EQUAL 64"), r.stdout + r.stderr
    r = subprocess.run([exe, os.path.join(DATA, "alice29.txt.compressed"), os.path.join(DATA, "alice29.txt")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("EQUAL 152089"), r.stdout + r.stderr
    r = subprocess.run([exe, os.path.join(DATA, "frewsxcv_06.compressed"), os.path.join(DATA, "64x")], capture_output=True, text=True)
    assert r.returncode == 3 and "ERR 23 Run length excceeded" in r.stdout, r.stdout + r.stderr
    # the streaming constructor (bro_reader_new_streaming): 2,000 bytes of input at a time
    r = subprocess.run([exe, "stream", os.path.join(DATA, "metablock_reset.compressed"), os.path.join(DATA, "metablock_reset"), "2000"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("EQUAL 912868"), r.stdout + r.stderr
    r = subprocess.run([exe, "stream", os.path.join(DATA, "frewsxcv_06.compressed"), os.path.join(DATA, "64x"), "7"], capture_output=True, text=True)
    assert r.returncode == 3 and "ERR 23 Run length excceeded" in r.stdout, r.stdout + r.stderr
    # brotli::decode_batch: several streams, no size hints
    names = ["64x", "alice29.txt", "quickfox_repeated", "random_org_10k.bin", "empty"]
    r = subprocess.run([exe, "batch"] + [os.path.join(DATA, n + ".compressed") for n in names] + [os.path.join(DATA, "frewsxcv_06.compressed")],
                       capture_output=True, text=True)
    want = ["0 %d" % os.path.getsize(os.path.join(DATA, n)) for n in names] + ["23 0"]
    assert r.returncode == 0 and r.stdout.split("\n")[:6] == want, r.stdout + r.stderr
    # brotli::decode_batch_multi_gpu (bro_mg_*): 240 streams over every GPU of the box, and over one
    for ngpus in ("0", "1"):
        r = subprocess.run([exe, "mg", ngpus] + [os.path.join(DATA, n + ".compressed") for n in names] + [os.path.join(DATA, "frewsxcv_06.compressed")],
                           capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.split("\n")[:6] == want, r.stdout + r.stderr


# ---- tools/corpus_driver.cpp: the reference's command-line driver (src/main.rs:49-70) ----

def _build_driver(tmp):
    from brotli_rs_b200 import _lib, build
    if not os.path.exists(_lib.library_path()):
        build.build()
    exe = os.path.join(tmp, "corpus_driver")
    libdir = os.path.dirname(_lib.library_path())
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "corpus_driver.cpp"),
                           "-o", exe, "-L" + libdir, "-lbrotli_b200", "-Wl,-rpath," + libdir])
    return exe


def test_corpus_driver_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build_driver(str(tmp_path))
    if not torch.cuda.is_available():
        r = subprocess.run([exe, DATA], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU decode path" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_corpus_driver_matches_oracle(tmp_path):
    """every *compressed file of the corpus, per file (the reference's call shape) and as one batch: the three lines the
    reference prints per file, with the oracle's length and error text"""
    from brotli_rs_b200 import _lib
    from oracle import oracle
    exe = _build_driver(str(tmp_path))
    names = sorted(fn for fn in os.listdir(DATA) if fn.endswith("compressed"))
    want = ""
    for fn in names:
        st, out = oracle.decode(open(os.path.join(DATA, fn), "rb").read())
        res = "Ok(%d)" % len(out) if st == 0 else 'Err(Custom { kind: InvalidData, error: "%s" })' % _lib.status_description(st)
        want += '"%s":\noutput length = %d\nres = %s\n===========\n\n' % (os.path.join(DATA, fn), len(out) if st == 0 else 0, res)
    assert len(names) >= 25
    for args in ([], ["--batch"]):
        r = subprocess.run([exe] + args + [DATA], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout == want, (args, r.stdout[:600], r.stderr[-300:])


# ---- the host-side C++ code on CPU, against tests/mock_abi.c (the oracle stands in for the device library) ----

def _build_mock(tmp):
    """-> directory holding a libbrotli_b200.so that is the mock (CPU test-suite only; see tests/mock_abi.c)"""
    from oracle import oracle
    oracle.lib()
    d = os.path.join(tmp, "mock")
    os.makedirs(d, exist_ok=True)
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests", "mock_abi.c"), "-o",
                           os.path.join(d, "libbrotli_b200.so"), "-L" + odir, "-l:liboracle.so", "-Wl,-rpath," + odir])
    return d


def _want_driver_output(names):
    from oracle import oracle
    want = ""
    for fn in names:
        st, out = oracle.decode(open(os.path.join(DATA, fn), "rb").read())
        res = "Ok(%d)" % len(out) if st == 0 else 'Err(Custom { kind: InvalidData, error: "%s" })' % oracle.lib().bro_oracle_status_description(st).decode()
        want += '"%s":\noutput length = %d\nres = %s\n===========\n\n' % (os.path.join(DATA, fn), len(out) if st == 0 else 0, res)
    return want


def test_host_code_against_mock_library(tmp_path):
    """brotli::Decompressor (whole-stream and streaming constructors), read_to_end, brotli::decode_batch and the corpus
    driver, compiled against the mock: the text the driver prints for the corpus (per file and --batch) is the
    reference's three lines per file with the oracle's lengths and error strings"""
    mock = _build_mock(str(tmp_path))
    inc = os.path.join(ROOT, "include")
    drv = os.path.join(str(tmp_path), "corpus_driver_mock")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", inc, os.path.join(ROOT, "tools", "corpus_driver.cpp"), "-o", drv,
                           "-L" + mock, "-lbrotli_b200", "-Wl,-rpath," + mock])
    names = sorted(fn for fn in os.listdir(DATA) if fn.endswith("compressed"))
    want = _want_driver_output(names)
    for args in ([], ["--batch"]):
        r = subprocess.run([drv] + args + [DATA], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout == want, (args, r.stdout[:400], r.stderr[-300:])
    r = subprocess.run([drv, os.path.join(DATA, "no_such_dir")], capture_output=True, text=True)
    assert r.returncode == 2
    # the twin program of this file: doc-test, error text, streaming constructor, batch
    src, exe = os.path.join(str(tmp_path), "twin_mock.cpp"), os.path.join(str(tmp_path), "twin_mock")
    open(src, "w").write(SRC)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", inc, src, "-o", exe, "-L" + mock, "-lbrotli_b200", "-Wl,-rpath," + mock])
    r = subprocess.run([exe, os.path.join(DATA, "64x.compressed"), os.path.join(DATA, "64x")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("EQUAL 64"), r.stdout + r.stderr
    r = subprocess.run([exe, os.path.join(DATA, "frewsxcv_06.compressed"), os.path.join(DATA, "64x")], capture_output=True, text=True)
    assert r.returncode == 3 and "ERR 23 Run length excceeded" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([exe, "stream", os.path.join(DATA, "metablock_reset.compressed"), os.path.join(DATA, "metablock_reset"), "2000"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("EQUAL 912868"), r.stdout + r.stderr
    bn = ["64x", "alice29.txt", "quickfox_repeated", "random_org_10k.bin", "empty"]
    r = subprocess.run([exe, "batch"] + [os.path.join(DATA, n + ".compressed") for n in bn] + [os.path.join(DATA, "frewsxcv_06.compressed")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split("\n")[:6] == ["0 %d" % os.path.getsize(os.path.join(DATA, n)) for n in bn] + ["23 0"], r.stdout + r.stderr
    r = subprocess.run([exe, "mg", "0"] + [os.path.join(DATA, n + ".compressed") for n in bn] + [os.path.join(DATA, "frewsxcv_06.compressed")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split("\n")[:6] == ["0 %d" % os.path.getsize(os.path.join(DATA, n)) for n in bn] + ["23 0"], r.stdout + r.stderr
