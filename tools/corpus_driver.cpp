// corpus_driver.cpp -- the reference's command-line driver (reference src/main.rs:49-70) on the B200 decoder.
//
// The reference's `main` walks data/, and for every file whose name ends in "compressed" prints the path, decodes it
// with Decompressor::new(File::open(path)).read_to_end(&mut input), and prints the output length and the io::Result.
// This driver prints the same three lines per file.  Default: one brotli::Decompressor per file (the reference's call
// shape, src/main.rs:61).  --batch: all files of the directory as ONE batch through brotli::decode_batch (no size
// hints), which is how a corpus is meant to be decoded on a GPU.  --suffix S overrides the name filter.
//
//   g++ -std=c++17 -O2 -I include tools/corpus_driver.cpp -o corpus_driver -L brotli_rs_b200/lib -lbrotli_b200 \
//       -Wl,-rpath,$PWD/brotli_rs_b200/lib
//   ./corpus_driver [--batch] [--suffix compressed] [DIR = data]
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <string>
#include <vector>

#include "brotli_b200.hpp"

namespace fs = std::filesystem;

static bool ends_with(const std::string& s, const std::string& suffix) {
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

// the three lines of src/main.rs:58-64; an error prints like Rust's io::Error built by src/lib.rs:2177
static void report(const std::string& path, size_t len, int status) {
    std::printf("\"%s\":\n", path.c_str());
    std::printf("output length = %zu\n", len);
    if (status == BRO_OK) std::printf("res = Ok(%zu)\n===========\n\n", len);
    else std::printf("res = Err(Custom { kind: InvalidData, error: \"%s\" })\n===========\n\n", bro_status_description(status));
}

int main(int argc, char** argv) {
    bool batch = false;
    std::string suffix = "compressed", dir = "data";
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--batch")) batch = true;
        else if (!std::strcmp(argv[i], "--suffix") && i + 1 < argc) suffix = argv[++i];
        else dir = argv[i];
    }
    std::error_code ec;
    if (!fs::is_directory(dir, ec)) {
        std::fprintf(stderr, "%s is not a directory\n", dir.c_str());
        return 2;
    }
    std::vector<std::string> paths;
    for (const auto& entry : fs::directory_iterator(dir))
        if (!entry.is_directory() && ends_with(entry.path().filename().string(), suffix)) paths.push_back(entry.path().string());
    std::sort(paths.begin(), paths.end());          // read_dir order is unspecified; sorted output can be compared
    bro_ctx* ctx = nullptr;
    if (bro_ctx_create(&ctx, -1) != BRO_OK) {
        std::fprintf(stderr, "no CUDA device / context (there is no CPU decode path)\n");
        return 3;
    }
    if (batch) {
        std::vector<std::vector<uint8_t>> streams;
        for (const auto& p : paths) {
            std::ifstream f(p, std::ios::binary);
            streams.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        }
        const auto items = brotli::decode_batch(ctx, streams);
        // the reference keeps the bytes decoded before an error in `input` (src/main.rs:63 prints their count); a batch
        // slot's content after an error is not part of the contract, so a failed stream reports length 0 here
        for (size_t i = 0; i < paths.size(); i++) report(paths[i], items[i].status == BRO_OK ? items[i].bytes.size() : 0, items[i].status);
    } else {
        for (const auto& p : paths) {
            std::ifstream f(p, std::ios::binary);
            brotli::Decompressor<brotli::IstreamReader> d{brotli::IstreamReader(f), ctx};
            std::vector<uint8_t> out;
            int status = BRO_OK;
            try { d.read_to_end(out); } catch (const brotli::Error& e) { status = e.status(); }
            report(p, status == BRO_OK ? out.size() : 0, status);
        }
    }
    bro_ctx_destroy(ctx);
    return 0;
}
