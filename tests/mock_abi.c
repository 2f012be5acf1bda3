/* mock_abi.c -- CPU TEST-SUITE ONLY: a stand-in for libbrotli_b200.so that implements the entry points the host-side C++
 * code uses (include/brotli_b200.hpp, tools/corpus_driver.cpp) with the parity oracle in place of the GPU, so that
 * `pytest -m "not gpu"` can run that host code -- the Read-struct twin, read_to_end, decode_batch, the corpus driver's
 * output -- where no GPU exists.  Built into tests/_build/mock/ by tests/test_cpp_twin.py; never shipped, never loaded
 * by the product package (whose library fails without a GPU).  Like tests/test_multirank_gloo.py, the oracle stands in
 * for the device here: what is tested is the host code around the decode, not the decoder. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/brotli_b200.h"

int bro_oracle_decode(const uint8_t* in, size_t in_len, uint8_t** out, size_t* out_len, int quirks);
void bro_oracle_free(void* p);
const char* bro_oracle_status_description(int st);

struct bro_ctx { int quirks; };
struct bro_reader {
    bro_read_cb cb; void* user; size_t chunk; int streaming;
    int decoded, status; uint8_t* out; size_t len, served;
};

int bro_ctx_create(bro_ctx** ctx, int device) { (void)device; *ctx = (bro_ctx*)calloc(1, sizeof(bro_ctx)); return *ctx ? BRO_OK : 104; }
void bro_ctx_destroy(bro_ctx* ctx) { free(ctx); }
const char* bro_status_description(int status) { return bro_oracle_status_description(status); }
void bro_free(void* p) { free(p); }

int bro_batch_decode_unsized_host(bro_ctx* ctx, const uint8_t* h_in, const uint64_t* h_in_off, uint32_t n, uint8_t** h_out,
                                  uint64_t* h_out_off, uint64_t* h_out_len, int32_t* h_status) {
    uint8_t** parts = (uint8_t**)calloc(n ? n : 1, sizeof(uint8_t*));
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) {
        size_t len = 0;
        h_status[i] = bro_oracle_decode(h_in + h_in_off[i], (size_t)(h_in_off[i + 1] - h_in_off[i]), &parts[i], &len, ctx ? ctx->quirks : 0);
        h_out_len[i] = len;
        h_out_off[i] = total;
        total += (len + 15u) & ~(uint64_t)15;
    }
    h_out_off[n] = total;
    *h_out = (uint8_t*)malloc(total ? total : 1);
    for (uint32_t i = 0; i < n; i++) {
        if (h_out_len[i]) memcpy(*h_out + h_out_off[i], parts[i], h_out_len[i]);
        bro_oracle_free(parts[i]);
    }
    free(parts);
    return BRO_OK;
}

static bro_reader* reader_new(bro_read_cb cb, void* user, size_t chunk, int streaming) {
    if (!cb) return NULL;
    bro_reader* r = (bro_reader*)calloc(1, sizeof(bro_reader));
    if (r) { r->cb = cb; r->user = user; r->chunk = chunk ? chunk : 65536; r->streaming = streaming; }
    return r;
}
bro_reader* bro_reader_new(bro_ctx* ctx, bro_read_cb cb, void* user) { (void)ctx; return reader_new(cb, user, 0, 0); }
bro_reader* bro_reader_new_streaming(bro_ctx* ctx, bro_read_cb cb, void* user, size_t in_chunk) { (void)ctx; return reader_new(cb, user, in_chunk, 1); }

intptr_t bro_reader_read(bro_reader* r, uint8_t* buf, size_t len) {
    if (!r->decoded) {
        uint8_t* in = NULL; size_t n = 0, cap = 0;
        for (;;) {
            if (n + r->chunk > cap) { cap = 2 * cap + r->chunk; in = (uint8_t*)realloc(in, cap); }
            intptr_t got = r->cb(r->user, in + n, r->chunk);
            if (got <= 0) break;
            n += (size_t)got;
        }
        r->status = bro_oracle_decode(in, n, &r->out, &r->len, 0);
        free(in);
        r->decoded = 1;
        /* the whole-stream reader reports an error with nothing in front of it; the streaming reader delivers what it
         * decoded first (the mock: everything the oracle produced before the error) */
        if (r->status != BRO_OK && !r->streaming) r->len = 0;
    }
    size_t left = r->len - r->served, k = len < left ? len : left;
    if (k) { memcpy(buf, r->out + r->served, k); r->served += k; return (intptr_t)k; }
    return r->status != BRO_OK ? -(intptr_t)r->status : 0;
}
int bro_reader_status(const bro_reader* r) { return r->status; }
void bro_reader_free(bro_reader* r) { if (r) { bro_oracle_free(r->out); free(r); } }

/* bro_mg_*: the mock has one "device" */
struct bro_mg { bro_ctx* ctx; };
int bro_mg_create(bro_mg** mg, int ngpus) {
    if (ngpus > 1) return 104;
    *mg = (bro_mg*)calloc(1, sizeof(bro_mg));
    return *mg ? bro_ctx_create(&(*mg)->ctx, 0) : 104;
}
void bro_mg_destroy(bro_mg* mg) { if (mg) { bro_ctx_destroy(mg->ctx); free(mg); } }
int bro_mg_device_count(const bro_mg* mg) { return mg ? 1 : 0; }
bro_ctx* bro_mg_ctx(bro_mg* mg, int k) { return (mg && k == 0) ? mg->ctx : NULL; }
int bro_mg_partition(const uint64_t* h_in_off, const uint64_t* h_out_off, uint32_t n, int ngpus, uint32_t* first) {
    (void)h_in_off; (void)h_out_off;
    for (int k = 0; k <= ngpus; k++) first[k] = k == 0 ? 0 : n;
    return BRO_OK;
}
int bro_mg_decode_host(bro_mg* mg, const uint8_t* h_in, const uint64_t* h_in_off, uint8_t* h_out, const uint64_t* h_out_off,
                       uint64_t* h_out_len, int32_t* h_status, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        uint8_t* part = NULL;
        size_t len = 0;
        h_status[i] = bro_oracle_decode(h_in + h_in_off[i], (size_t)(h_in_off[i + 1] - h_in_off[i]), &part, &len, mg->ctx->quirks);
        if (h_status[i] == BRO_OK && len > h_out_off[i + 1] - h_out_off[i]) h_status[i] = BRO_OUTPUT_TOO_SMALL;
        h_out_len[i] = h_status[i] == BRO_OK ? len : 0;
        if (h_status[i] == BRO_OK && len) memcpy(h_out + h_out_off[i], part, len);
        bro_oracle_free(part);
    }
    return BRO_OK;
}
