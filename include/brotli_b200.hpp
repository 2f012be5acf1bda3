// brotli_b200.hpp -- C++ twin of the reference's one public type, brotli::Decompressor<R: Read>
// (reference src/lib.rs:378-410 `pub struct Decompressor<R: Read>` / `new`, and src/lib.rs:2173-2193 `impl Read`),
// written above the C ABI of brotli_b200.h.  The reference is Rust; no Rust toolchain exists in the build image, so
// the host side above the ABI is C++ (header only).  INTEGRATION.md carries the Rust shim a maintainer would add.
//
//   std::ifstream f("data/64x.compressed", std::ios::binary);
//   brotli::Decompressor<brotli::IstreamReader> d{brotli::IstreamReader(f)};
//   std::vector<uint8_t> out;
//   d.read_to_end(out);                       // == Decompressor::new(f).read_to_end(&mut out)
//
// R is any type with `size_t read(uint8_t* buf, size_t cap)` returning 0 at end of input (std::io::Read::read).
// Errors surface like the reference's io::Error::new(ErrorKind::InvalidData, description): brotli::Error carries the
// status (DecompressorError number, src/lib.rs:294-319) and what() is the reference's description string.
#pragma once
#include <cstdint>
#include <istream>
#include <stdexcept>
#include <utility>
#include <vector>

#include "brotli_b200.h"

namespace brotli {

class Error : public std::runtime_error {
public:
    explicit Error(int status) : std::runtime_error(bro_status_description(status)), status_(status) {}
    int status() const { return status_; }
private:
    int status_;
};

// Read adapter over std::istream (File / Cursor in the reference's tests).
class IstreamReader {
public:
    explicit IstreamReader(std::istream& s) : s_(&s) {}
    size_t read(uint8_t* buf, size_t cap) {
        s_->read(reinterpret_cast<char*>(buf), static_cast<std::streamsize>(cap));
        return static_cast<size_t>(s_->gcount());
    }
private:
    std::istream* s_;
};

// Read adapter over a byte slice (`&[u8]` in the reference's tests).
class SliceReader {
public:
    SliceReader(const uint8_t* p, size_t n) : p_(p), n_(n) {}
    size_t read(uint8_t* buf, size_t cap) {
        size_t k = cap < n_ ? cap : n_;
        for (size_t i = 0; i < k; i++) buf[i] = p_[i];
        p_ += k; n_ -= k;
        return k;
    }
private:
    const uint8_t* p_;
    size_t n_;
};

template <class R>
class Decompressor {
public:
    // Decompressor::new(r): infallible, performs no I/O (src/lib.rs:398-410).  `ctx` may be shared between readers
    // used from the same thread; with nullptr the reader creates its own context on the current CUDA device.
    explicit Decompressor(R r, bro_ctx* ctx = nullptr) : r_(std::move(r)), h_(bro_reader_new(ctx, &trampoline, this)) {
        if (!h_) throw std::bad_alloc();
    }
    // The same with the reference's memory behaviour (bro_reader_new_streaming): `r` is asked for in_chunk bytes at a
    // time (0 = 1 MiB) and decoded meta-block by meta-block; buffers are bounded by a meta-block plus a window.
    struct Streaming { size_t in_chunk; };
    Decompressor(R r, Streaming s, bro_ctx* ctx = nullptr)
        : r_(std::move(r)), h_(bro_reader_new_streaming(ctx, &trampoline, this, s.in_chunk)) {
        if (!h_) throw std::bad_alloc();
    }
    Decompressor(const Decompressor&) = delete;
    Decompressor& operator=(const Decompressor&) = delete;
    ~Decompressor() { bro_reader_free(h_); }

    // Read::read (src/lib.rs:2174-2192): fills buf while data is available; 0 at end of stream; throws on InvalidData.
    size_t read(uint8_t* buf, size_t len) {
        intptr_t n = bro_reader_read(h_, buf, len);
        if (n < 0) throw Error(static_cast<int>(-n));
        return static_cast<size_t>(n);
    }

    // Read::read_to_end, the call shape of every reference test, bench and the CLI (src/main.rs:61).
    size_t read_to_end(std::vector<uint8_t>& out) {
        size_t total = 0;
        uint8_t chunk[1 << 16];
        for (;;) {
            size_t n = read(chunk, sizeof(chunk));
            if (n == 0) return total;
            out.insert(out.end(), chunk, chunk + n);
            total += n;
        }
    }

    int status() const { return bro_reader_status(h_); }

private:
    static intptr_t trampoline(void* user, uint8_t* buf, size_t cap) {
        try {
            return static_cast<intptr_t>(static_cast<Decompressor*>(user)->r_.read(buf, cap));
        } catch (...) {
            return -1;   // an input I/O error ends the stream: the reference folds it into UnexpectedEOF
        }
    }
    R r_;
    bro_reader* h_;
};

// Many streams at once (no counterpart in the reference, which has one stream per Decompressor): the batch entry
// point a host program would use, without size hints (bro_batch_decode_unsized_host).  Element i = {status, bytes};
// status is the DecompressorError number of src/lib.rs:294-319 (0 = valid stream).
struct BatchItem {
    int status;
    std::vector<uint8_t> bytes;
};

inline std::vector<BatchItem> decode_batch(bro_ctx* ctx, const std::vector<std::vector<uint8_t>>& streams) {
    const uint32_t n = static_cast<uint32_t>(streams.size());
    std::vector<uint64_t> in_off(n + 1, 0), out_off(n + 1, 0), out_len(n, 0);
    std::vector<int32_t> status(n, 0);
    for (uint32_t i = 0; i < n; i++) in_off[i + 1] = in_off[i] + streams[i].size();
    std::vector<uint8_t> in(in_off[n] ? in_off[n] : 1);
    for (uint32_t i = 0; i < n; i++)
        for (size_t k = 0; k < streams[i].size(); k++) in[in_off[i] + k] = streams[i][k];
    uint8_t* out = nullptr;
    const int st = bro_batch_decode_unsized_host(ctx, in.data(), in_off.data(), n, &out, out_off.data(), out_len.data(), status.data());
    if (st != BRO_OK) throw Error(st);
    std::vector<BatchItem> res(n);
    for (uint32_t i = 0; i < n; i++) {
        res[i].status = status[i];
        if (out_len[i]) res[i].bytes.assign(out + out_off[i], out + out_off[i] + out_len[i]);
    }
    bro_free(out);
    return res;
}

// One batch over several GPUs of the box from this process (bro_mg_*): streams and the capacity of every stream's slot in,
// status and bytes out.  ngpus <= 0: every visible device.
inline std::vector<BatchItem> decode_batch_multi_gpu(const std::vector<std::vector<uint8_t>>& streams, const std::vector<size_t>& capacities,
                                                     int ngpus = 0) {
    const uint32_t n = static_cast<uint32_t>(streams.size());
    std::vector<uint64_t> in_off(n + 1, 0), out_off(n + 1, 0), out_len(n, 0);
    std::vector<int32_t> status(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        in_off[i + 1] = in_off[i] + streams[i].size();
        out_off[i + 1] = out_off[i] + ((capacities[i] + 15) & ~static_cast<size_t>(15));      // 16-byte aligned slot starts
    }
    std::vector<uint8_t> in(in_off[n] + 16), out(out_off[n] ? out_off[n] : 1);
    for (uint32_t i = 0; i < n; i++)
        for (size_t k = 0; k < streams[i].size(); k++) in[in_off[i] + k] = streams[i][k];
    bro_mg* mg = nullptr;
    int st = bro_mg_create(&mg, ngpus);
    if (st != BRO_OK) throw Error(st);
    st = bro_mg_decode_host(mg, in.data(), in_off.data(), out.data(), out_off.data(), out_len.data(), status.data(), n);
    bro_mg_destroy(mg);
    if (st != BRO_OK) throw Error(st);
    std::vector<BatchItem> res(n);
    for (uint32_t i = 0; i < n; i++) {
        res[i].status = status[i];
        if (status[i] == BRO_OK && out_len[i]) res[i].bytes.assign(out.begin() + out_off[i], out.begin() + out_off[i] + out_len[i]);
    }
    return res;
}

}  // namespace brotli
