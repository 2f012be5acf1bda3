// bro_warpsim.cpp -- 32-LANE host simulation of the fused warp-per-stream decoder -- CPU TEST-SUITE ONLY.
// Built into tests/_build/libbro_warpsim.so by tests/warpsim.py; never part of libbrotli_b200.so.
//
// bro_hostsim.cpp runs bro_decoder_core.h with a 1-lane "warp" (the thread-per-stream form).  The code the fused kernel
// (bro_kernels.cu: bro_decode_warp_kernel) and the resume kernel actually run -- BRO_W = 32: the lane-parallel table
// build, the shuffle-fed bit window, the lane-parallel literal rounds, bro_lz_copy / bro_copy_far / bro_dict_word --
// is compiled HERE, unchanged, with g++: the 32 lanes are 32 fibers (one stack each) on one OS thread, and the handful
// of warp intrinsics the header uses (__shfl_sync, __match_any_sync, __all_sync, __syncwarp) are rendezvous points
// between them.  Between two rendezvous a lane runs alone, and the order in which the lanes run is the caller's choice
// (ascending, descending, a seeded shuffle re-drawn at every rendezvous): a store that another lane reads without a
// __syncwarp in between -- which the hardware may or may not order -- becomes a deterministic wrong answer in one of
// the orders.  The simulation also checks what the kernel relies on: every lane arrives at the SAME intrinsic with the
// same mask (no divergent collective), no lane leaves while others wait for it, and all 32 lanes return the same status
// and output length (the decoder state is warp-uniform).
//
// What it does not model: memory-access timing, the read-only (__ldg) path, shared-memory address spaces.
#if !defined(__x86_64__)
#error "bro_warpsim.cpp: the fiber switch below is x86-64 only (tests/warpsim.py skips the suite elsewhere)"
#endif
#define BRO_WARPSIM 1
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bro_warpsim.h"

// bro_kernels.cu itself (it includes bro_decoder_core.h in the 32-lane form): the fused kernel's work-queue loop, retry pass and
// hand-out order are run by bro_warpsim_fused_launch below; the per-stream entry points further down call the decoder directly
static uint8_t ws_dynamic_smem[96 * 1024] __attribute__((aligned(128)));
#include "bro_kernels.cu"
#include "bro_kernels_resume.cu"

extern "C" const uint8_t bro_dictionary_blob[];

// ------------------------------------------------------------------------------------------------------
// what one warp of bro_decode_warp_kernel / bro_decode_resume_kernel does for one stream (bro_kernels.cu,
// bro_kernels_resume.cu), lane by lane
// ------------------------------------------------------------------------------------------------------
struct WsJob {
    const uint8_t* in; size_t in_len;
    uint8_t* out; size_t cap;
    int quirks, latency;
    BroScratch* scratch;
    uint16_t* root10;
    uint16_t* arena;
    BroResume* ck;            // != 0: the resume kernel
    uint64_t r_in_off[2], r_out_off[2], r_out_len;      // its launch arguments for a batch of one
    int32_t r_status;
    uint32_t r_counter;
};

WS_NO_TSAN static WsLane* ws_me() { return &g_ws->lane[g_ws->cur]; }
WS_NO_TSAN static void ws_lane_result(WsLane* me, int st, uint32_t pos) { me->ret_status = st; me->ret_pos = pos; }

static void ws_lane_body(void* arg) {
    WsJob* j = (WsJob*)arg;
    WsLane* me = ws_me();
    BroDec d;
    memset(&d, 0xa5, sizeof(d));           // whatever the kernel does not set is garbage there too
    d.sc = j->scratch;
    d.root10 = j->latency && !j->ck ? j->root10 : (uint16_t*)0;
    d.arena = j->arena;
    d.arena_cap = BRO_ARENA_U16_MAX;
    d.arena_base = 0;
    d.dict = bro_dictionary_blob;
    d.out = j->out;
    d.cap = j->cap > BRO_MAX_SLOT ? (uint32_t)BRO_MAX_SLOT : (uint32_t)j->cap;
    d.quirk_spec = j->quirks;
    int st;
    if (j->ck) {
        // the resume kernel itself (bro_kernels_resume.cu), a batch of one stream
        BroLaunch p;
        memset(&p, 0, sizeof(p));
        p.in = j->in; p.in_off = j->r_in_off; p.out = j->out; p.out_off = j->r_out_off; p.out_len = &j->r_out_len; p.status = &j->r_status;
        p.n = 1; p.arena = j->arena; p.dict = bro_dictionary_blob; p.counter = &j->r_counter; p.quirk_spec = j->quirks; p.resume = j->ck;
        bro_decode_resume_kernel(p);
        __syncwarp();                      // (lane 0's results are read below by every lane)
        st = j->r_status;
        d.pos = (uint32_t)j->r_out_len;
    } else {
        d.pos = 0;
        d.p1 = 0; d.p2 = 0;
        d.d0 = 4; d.d1 = 11; d.d2 = 15; d.d3 = 16;
        bro_bits_init(d.in, j->in, j->in + j->in_len);
        st = bro_decode_stream(d);
        bro_syncwarp();
    }
    ws_lane_result(me, st, d.pos);
}

static uint64_t g_last_rendezvous;

// in / out: the caller's buffers; the simulation decodes from and into padded copies (the kernels read whole aligned
// words and 16-byte granules around what they need, inside allocations the product pads).
// order: 0 ascending, 1 descending, 2 shuffled (seed).  *sim_err: WS_ERR_* (0 = the warp behaved).
static size_t g_guard_in_slack = 16, g_guard_out_slack = 0;
static unsigned g_in_mis, g_out_mis;      // address of the stream's first byte mod 128, of the slot's first byte mod 16

static int ws_decode(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap, size_t* out_len, int quirks, int latency,
                     int order, uint64_t seed, BroResume* ck, int* sim_err) {
    enum { PAD = 256 };
    uint8_t* inb = (uint8_t*)malloc(in_len + 2 * PAD + 128);
    uint8_t* in_al = (uint8_t*)(((uintptr_t)inb + PAD + 127) & ~(uintptr_t)127) + (g_in_mis & 127u);
    memset(inb, 0xee, in_len + 2 * PAD + 128);
    memcpy(in_al, in, in_len);
    uint8_t* outb = (uint8_t*)malloc(cap + 2 * PAD + 16);
    uint8_t* out_al = (uint8_t*)(((uintptr_t)outb + PAD + 15) & ~(uintptr_t)15) + (g_out_mis & 15u);
    memset(outb, 0xdd, cap + 2 * PAD + 16);
    if (ws_guard_mode) {
        // the buffers end (or begin) at unmapped pages: the compressed stream with the 16 readable bytes behind it that
        // include/brotli_b200.h asks of a caller, the output slot with g_guard_out_slack
        in_al = ws_guard_alloc(in_len, g_guard_in_slack, 0, 0xee);
        memcpy(in_al, in, in_len);
        out_al = ws_guard_alloc(cap, g_guard_out_slack, 0, 0xdd);
    }
    if (ck) memcpy(out_al, out, ck->pos <= cap ? ck->pos : cap);         // the history
    WsJob j;
    j.in = in_al; j.in_len = in_len; j.out = out_al; j.cap = cap; j.quirks = quirks; j.latency = latency; j.ck = ck;
    j.r_in_off[0] = 0; j.r_in_off[1] = in_len; j.r_out_off[0] = 0; j.r_out_off[1] = cap; j.r_out_len = 0; j.r_status = -1; j.r_counter = 0;
    j.scratch = (BroScratch*)aligned_alloc(16, (sizeof(BroScratch) + 15u) & ~(size_t)15);
    memset(j.scratch, 0xcc, sizeof(BroScratch));
    j.root10 = (uint16_t*)malloc(2048);
    memset(j.root10, 0xcc, 2048);
    j.arena = (uint16_t*)aligned_alloc(128, ((2u * (size_t)BRO_ARENA_U16_MAX) + 127u) & ~(size_t)127);
    WsWarp* w = (WsWarp*)malloc(sizeof(WsWarp));
    const int err = ws_run(w, ws_lane_body, &j, order, seed);
    const int st = w->lane[0].ret_status;
    size_t n = w->lane[0].ret_pos;
    g_last_rendezvous = w->rendezvous;
    *sim_err = err;
    if (err) n = 0;
    // nothing outside the slot may have been written
    bool clobber = false;
    if (!ws_guard_mode) {
        for (uint8_t* p = outb; p < out_al && !clobber; p++) clobber = *p != 0xdd;
        for (uint8_t* p = out_al + cap; p < outb + cap + 2 * PAD + 16 && !clobber; p++) clobber = *p != 0xdd;
    } else if (ws_guard_mode == 1) for (size_t k = 0; k < g_guard_out_slack; k++) clobber |= out_al[cap + k] != 0xdd;
    if (clobber && !*sim_err) *sim_err = 100;
    if (n > cap) n = cap;
    // (the resume kernel: the whole slot goes back -- a call that failed inside a meta-block reports the position it had reached,
    // which may lie beyond the slot when the stores that did not fit were skipped; only what lies in front of ck->pos is final)
    memcpy(out, out_al, ck ? (err ? 0 : cap) : n);
    *out_len = n;
    free(w); free(j.arena); free(j.root10); free(j.scratch); free(outb); free(inb);
    return st;
}

extern "C" int bro_warpsim_decode(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap, size_t* out_len, int quirks,
                                  int latency, int order, uint64_t seed, int* sim_err) {
    return ws_decode(in, in_len, out, cap, out_len, quirks, latency, order, seed, 0, sim_err);
}

// one call of the resume kernel for one stream: out[0 .. ck->pos) holds the history on entry; on return the whole slot
// is copied back (only the bytes in front of the new ck->pos are final) and *out_len = bytes in the slot when the call stopped
extern "C" int bro_warpsim_decode_resume(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap, size_t* out_len, int quirks,
                                         int order, uint64_t seed, BroResume* ck, int* sim_err) {
    return ws_decode(in, in_len, out, cap, out_len, quirks, 0, order, seed, ck, sim_err);
}

// where the next decodes place the compressed stream (mod 128: the warp loads 128-byte chunks) and the output slot (mod 16:
// the copies store 16-byte vectors) -- on the device both are wherever the batch's offsets put them
extern "C" void bro_warpsim_set_alignment(unsigned in_mis, unsigned out_mis) { g_in_mis = in_mis; g_out_mis = out_mis; }
extern "C" uint64_t bro_warpsim_last_rendezvous() { return g_last_rendezvous; }
extern "C" unsigned bro_warpsim_resume_bytes() { return (unsigned)sizeof(BroResume); }

// barrier statistics: hits[line] = executions of the bro_syncwarp() on that line of bro_decoder_core.h since the last reset
extern "C" void bro_warpsim_sync_hits(uint64_t* hits, int n, int reset) {
    for (int i = 0; i < n && i < WS_MAX_LINE; i++) hits[i] = g_sync_hits[i];
    if (reset) memset(g_sync_hits, 0, sizeof(g_sync_hits));
}
extern "C" void bro_warpsim_drop_sync(int line) { g_sync_drop_line = line; }

// ------------------------------------------------------------------------------------------------------
// one launch of bro_decode_warp_kernel over a batch: one CTA, of which warp 0 runs (the others would only take other streams)
// ------------------------------------------------------------------------------------------------------
struct WsFusedJob { BroLaunch p; int latency; };
static void ws_fused_lane(void* arg) {
    WsFusedJob* j = (WsFusedJob*)arg;
    if (j->latency) bro_decode_warp_kernel<BRO_WARPS_PER_CTA, BRO_MIN_BLOCKS_LATENCY>(j->p);
    else bro_decode_warp_kernel<BRO_WARPS_PER_CTA, BRO_MIN_BLOCKS>(j->p);
    ws_lane_result(ws_me(), 0, 0);
}
// Buffers as in BroLaunch, owned by the caller (addressable a few bytes either side of the batch).  retry_mode: only the streams
// whose status[] holds one of the hand-over codes are decoded (the second pass of the two-phase path); order: hand-out order or NULL.
extern "C" int bro_warpsim_fused_launch(const uint8_t* in, const uint64_t* in_off, uint8_t* out, const uint64_t* out_off, uint64_t* out_len,
                                        int32_t* status, uint32_t n, const uint32_t* order, int retry_mode, int latency, int quirks,
                                        int lane_order, uint64_t seed, int nthreads) {
    WsFusedJob j;
    memset(&j, 0, sizeof(j));
    uint32_t counter = 0, retry = 0;
    for (uint32_t i = 0; i < n; i++) retry += retry_mode && BRO_ST_IS_RETRY(status[i]);
    if (nthreads < 32) nthreads = 32;
    if (nthreads > 32 * BRO_WARPS_PER_CTA) abort();
    uint16_t* arena = (uint16_t*)aligned_alloc(128, ((2u * (size_t)BRO_ARENA_U16_MAX * BRO_WARPS_PER_CTA) + 127u) & ~(size_t)127);
    j.p.in = in; j.p.in_off = in_off; j.p.out = out; j.p.out_off = out_off; j.p.out_len = out_len; j.p.status = status; j.p.n = n;
    j.p.arena = arena; j.p.dict = bro_dictionary_blob; j.p.counter = &counter; j.p.order = order; j.p.retry_count = &retry;
    j.p.retry_mode = retry_mode; j.p.quirk_spec = quirks;
    j.latency = latency;
    memset(ws_dynamic_smem, 0xcc, sizeof(ws_dynamic_smem));
    WsWarp* w = (WsWarp*)malloc(sizeof(WsWarp));
    const int err = ws_run(w, ws_fused_lane, &j, lane_order, seed, nthreads);
    g_last_rendezvous = w->rendezvous;
    free(w); free(arena);
    return err;
}

#if defined(BRO_WARPSIM_MAIN)
// warpsim_tsan <latency 0|1> <order 0|1|2> <seed> <quirks 0|1> <slack> file...: every file is one compressed stream, decoded
// into a slot of (bytes the oracle produced, passed as "file:size") + slack.  Prints "name status out_len fnv1a64(out)" per stream;
// ThreadSanitizer's reports go to stderr and make the exit code non-zero.
int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s latency order seed quirks slack file:size...\n", argv[0]); return 2; }
    const int latency = atoi(argv[1]), order = atoi(argv[2]), quirks = atoi(argv[4]);
    const uint64_t seed = strtoull(argv[3], 0, 10);
    const size_t slack = strtoull(argv[5], 0, 10);
    int bad = 0;
    if (getenv("BRO_WS_ALIGN")) { unsigned a = 0, b = 0; sscanf(getenv("BRO_WS_ALIGN"), "%u,%u", &a, &b); bro_warpsim_set_alignment(a, b); }
    if (getenv("BRO_WS_GUARD")) {        // "back,<in slack>,<out slack>" or "front"
        const char* g = getenv("BRO_WS_GUARD");
        unsigned a = 16, b = 0;
        ws_guard_mode = g[0] == 'f' ? 2 : 1;
        if (ws_guard_mode == 1) sscanf(g, "back,%u,%u", &a, &b);
        g_guard_in_slack = a; g_guard_out_slack = b;
        ws_guard_install();
    }
    if (getenv("BRO_WS_DROP_SYNC")) g_sync_drop_line = atoi(getenv("BRO_WS_DROP_SYNC"));      // mutation: the report must name it
    if (getenv("BRO_WS_BATCH")) {
        // BRO_WS_BATCH=<threads>: the files as ONE batch through bro_decode_warp_kernel itself, a CTA of threads / 32 warps that take
        // streams from the work queue side by side (each with its own scratch block and table arena: nothing may be shared)
        const int threads = atoi(getenv("BRO_WS_BATCH"));
        const uint32_t n = (uint32_t)(argc - 6);
        uint64_t* in_off = (uint64_t*)calloc(2u * (n + 1u), sizeof(uint64_t));
        uint64_t* out_off = in_off + (n + 1u);
        const size_t PADB = 256;
        uint8_t* in = (uint8_t*)calloc(2 * PADB, 1);
        for (uint32_t i = 0; i < n; i++) {
            char* colon = strrchr(argv[6 + i], ':');
            if (!colon) return 2;
            *colon = 0;
            out_off[i + 1] = out_off[i] + strtoull(colon + 1, 0, 10) + slack;
            FILE* f = fopen(argv[6 + i], "rb");
            if (!f) { perror(argv[6 + i]); return 2; }
            fseek(f, 0, SEEK_END);
            const size_t len = (size_t)ftell(f);
            fseek(f, 0, SEEK_SET);
            in = (uint8_t*)realloc(in, PADB + (size_t)in_off[i] + len + PADB);
            if (fread(in + PADB + in_off[i], 1, len, f) != len) return 2;
            fclose(f);
            in_off[i + 1] = in_off[i] + len;
        }
        memset(in + PADB + in_off[n], 0, PADB);
        uint8_t* out = (uint8_t*)calloc((size_t)out_off[n] + 2 * PADB, 1);
        uint64_t* out_len = (uint64_t*)calloc(n + 1u, sizeof(uint64_t));
        int32_t* status = (int32_t*)calloc(n + 1u, sizeof(int32_t));
        const int err = bro_warpsim_fused_launch(in + PADB, in_off, out + PADB, out_off, out_len, status, n, 0, 0, latency, quirks, order, seed, threads);
        for (uint32_t i = 0; i < n; i++) {
            uint64_t h = 1469598103934665603ull;
            const uint64_t cap = out_off[i + 1] - out_off[i], len = out_len[i] < cap ? out_len[i] : cap;
            for (uint64_t k = 0; k < len; k++) h = (h ^ out[PADB + out_off[i] + k]) * 1099511628211ull;
            printf("%s %d %llu %016llx %d\n", argv[6 + i], status[i], (unsigned long long)len, (unsigned long long)h, err);
        }
        return err ? 3 : 0;
    }
    for (int a = 6; a < argc; a++) {
        char* colon = strrchr(argv[a], ':');
        if (!colon) return 2;
        *colon = 0;
        const size_t cap = strtoull(colon + 1, 0, 10) + slack;
        FILE* f = fopen(argv[a], "rb");
        if (!f) { perror(argv[a]); return 2; }
        fseek(f, 0, SEEK_END);
        const size_t n = (size_t)ftell(f);
        fseek(f, 0, SEEK_SET);
        uint8_t* in = (uint8_t*)malloc(n + 1);
        if (fread(in, 1, n, f) != n) return 2;
        fclose(f);
        uint8_t* out = (uint8_t*)malloc(cap + 1);
        size_t out_len = 0;
        int sim_err = 0;
        const int st = bro_warpsim_decode(in, n, out, cap, &out_len, quirks, latency, order, seed, &sim_err);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < out_len; i++) h = (h ^ out[i]) * 1099511628211ull;
        printf("%s %d %zu %016llx %d\n", argv[a], st, out_len, (unsigned long long)h, sim_err);
        bad |= sim_err;
        free(out); free(in);
    }
    return bad ? 3 : 0;
}
#endif
