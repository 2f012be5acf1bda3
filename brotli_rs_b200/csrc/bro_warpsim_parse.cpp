// bro_warpsim_parse.cpp -- 32-LANE host simulation of the PARSE KERNEL (phase one of the two-phase path) -- CPU TEST-SUITE ONLY.
// Built into tests/_build/libbro_warpsim_parse.so by tests/warpsim.py; never part of libbrotli_b200.so.
//
// bro_kernels_parse.cu is compiled HERE with g++ -- the kernel function itself, 32 streams to a warp: the boundary protocol
// (lanes wait for each other at meta-block boundaries, with patience; finished lanes report, announce their stream in the
// completion queue and pull the next ones with one atomic of a leader; `lanes` < 32 for small batches), the lockstep rounds of
// bro_parse.h with lanes in and out of immediate mode side by side, and the lane-interleaved shared-memory blocks with their
// real window addresses.  tests/hostsim.py runs the same per-lane code for ONE stream at a time; what only 32 lanes together can
// show -- a lane reaching into a neighbour's block, a warp-wide vote taken in the wrong place, a stream lost or decoded twice by
// the hand-out -- shows here.  The device's inline PTX has C++ twins in bro_decoder_core.h (BRO_WARPSIM): shared-memory
// accesses by window address, and the compressed words' ring, whose asynchronous copies are performed as LATE as the kernel's
// cp.async.wait_group allows -- a wait that is one group short reads a stale word here, every time.
#define BRO_WARPSIM 1
#include "bro_warpsim.h"

// the CTA: shared memory for two warps (the second one only fills its part of the insert/copy table: a launch has >= 64 threads)
static uint8_t ws_dynamic_smem[8 * 32 * 1024] __attribute__((aligned(128)));
static unsigned ws_tid_base;
struct WsTidB { unsigned x; };
#undef threadIdx
#define threadIdx (WsTidB{ws_tid() + ws_tid_base})

// the ring's asynchronous copies (global -> shared, 4 bytes, one commit group each), per lane, oldest first
struct WsRingCopy { uint32_t dst; const void* src; };
static WsRingCopy g_ring[WS_MAX_THREADS][64];
static unsigned g_ring_head[WS_MAX_THREADS], g_ring_n[WS_MAX_THREADS];
static int g_ring_late = 1;          // 1: a copy lands when a wait forces it (the latest the hardware may); 0: at once
static inline void ws_ring_land(const WsRingCopy& c) { *(uint32_t*)ws_smem_ptr(c.dst) = *(const uint32_t*)c.src; }
static unsigned g_ring_slack = 0;     // mutation: every wait lets this many more groups stay in flight than the kernel asks for
static inline void ws_ring_wait_group(unsigned keep) {
    const unsigned l = ws_tid();
    if (keep) keep += g_ring_slack;
    while (g_ring_n[l] > keep) { ws_ring_land(g_ring[l][g_ring_head[l]]); g_ring_head[l] = (g_ring_head[l] + 1u) & 63u; g_ring_n[l]--; }
}
static inline void ws_ring_issue(uint32_t dst, const void* src) {
    const unsigned l = ws_tid();
    WsRingCopy c; c.dst = dst; c.src = src;
    if (!g_ring_late) { ws_ring_land(c); return; }
    if (g_ring_n[l] >= 64u) abort();
    g_ring[l][(g_ring_head[l] + g_ring_n[l]) & 63u] = c;
    g_ring_n[l]++;
}

#include "bro_kernels_parse.cu"

extern "C" const uint8_t bro_dictionary_blob[];

static void ws_parse_lane(void* arg) { bro_parse_kernel(*(BroLaunch*)arg); }

static uint64_t g_parse_rendezvous;

// One launch of the parse kernel over a batch, one warp.  Buffers as in BroLaunch, owned by the caller: `in` and `out` must be
// addressable a few bytes either side of the batch (the kernel reads whole aligned words), `rec` 16-byte aligned with
// rec_total records.  order: the hand-out order (NULL = identity); lanes: streams the warp holds at a time (1..32).
// done_q (n entries) receives the completion order -- what the copy kernel consumes.  Returns the simulation's verdict.
extern "C" int bro_warpsim_parse_launch(const uint8_t* in, const uint64_t* in_off, uint8_t* out, const uint64_t* out_off, uint64_t* out_len,
                                        int32_t* status, uint32_t* nrec, uint32_t* rec_words, uint64_t rec_total, uint32_t n, const uint32_t* order,
                                        uint32_t lanes, int quirks, int sizing, int lane_order, uint64_t seed, int ring_late, uint32_t* done_q,
                                        uint32_t* retry_count, int nthreads) {
    BroLaunch p;
    memset(&p, 0, sizeof(p));
    uint32_t counter = 0, done_tail = 0, retry = 0;
    uint16_t* arena = (uint16_t*)aligned_alloc(128, (2u * (size_t)BRO_THREAD_ARENA_STRIDE_U16 * WS_MAX_THREADS + 127u) & ~(size_t)127);
    p.in = in; p.in_off = in_off; p.out = out; p.out_off = out_off; p.out_len = out_len; p.status = status;
    p.arena = arena; p.dict = bro_dictionary_blob; p.counter = &counter; p.order = order; p.retry_count = &retry;
    p.quirk_spec = quirks; p.rec = (BroRec*)rec_words; p.rec_total = rec_total; p.nrec = nrec; p.done_q = done_q; p.done_tail = &done_tail;
    p.gate = 0; p.sizing = sizing; p.lanes = lanes;
    for (uint32_t i = 0; i < n; i++) { status[i] = -12345; if (done_q) done_q[i] = 0xffffffffu; }
    g_ring_late = ring_late;
    memset(g_ring_n, 0, sizeof(g_ring_n));
    memset(ws_dynamic_smem, 0xcc, sizeof(ws_dynamic_smem));
    WsWarp* w = (WsWarp*)malloc(sizeof(WsWarp));
    int err = 0;
    if (nthreads <= 32) {
        // one warp decodes.  (A launch has >= 64 threads -- the first 48 fill the insert/copy table: the CTA's second warp runs
        // first, finds no stream (n = 0 for it) and leaves after its share of the table)
        p.n = 0;
        ws_tid_base = 32u;
        err = ws_run(w, ws_parse_lane, &p, 0, 1);
        p.n = n;
        counter = 0;
        ws_tid_base = 0u;
        memset(g_ring_n, 0, sizeof(g_ring_n));
        if (!err) err = ws_run(w, ws_parse_lane, &p, lane_order, seed);
    } else {
        // nthreads / 32 warps side by side: they compete for the streams and interleave their entries in the completion queue
        p.n = n;
        ws_tid_base = 0u;
        err = ws_run(w, ws_parse_lane, &p, lane_order, seed, nthreads);
    }
    g_parse_rendezvous = w->rendezvous;
    // every stream reported exactly once, and announced exactly once
    if (!err) {
        for (uint32_t i = 0; i < n && !err; i++) if (status[i] == -12345) err = 102;
        if (!err && done_q && done_tail != n) { err = 103; if (getenv("BRO_WS_DEBUG")) fprintf(stderr, "done_tail %u n %u counter %u\n", done_tail, n, counter); }
        if (done_q && !err) {
            uint8_t* seen = (uint8_t*)calloc(n + 1u, 1);
            for (uint32_t i = 0; i < n && !err; i++) { if (done_q[i] >= n || seen[done_q[i]]) err = 103; else seen[done_q[i]] = 1; }
            free(seen);
        }
    }
    if (retry_count) *retry_count = retry;
    free(w); free(arena);
    return err;
}
// ------------------------------------------------------------------------------------------------------
// the size-class ordering kernels (what bro_order_launch launches), CTA by CTA: order[] <- stream indices by size class, largest
// first; gate[0] = longest compressed stream, gate[1] = AUTO's verdict
// ------------------------------------------------------------------------------------------------------
struct WsOrderJob { const uint64_t* in_off; uint32_t n; uint32_t* scratch; uint32_t* order; uint32_t* gate; };
static void ws_order_hist(void* a) { WsOrderJob* j = (WsOrderJob*)a; bro_order_hist_kernel(j->in_off, j->n, j->scratch, j->gate); }
static void ws_order_scan(void* a) { WsOrderJob* j = (WsOrderJob*)a; bro_order_scan_kernel(j->scratch, j->scratch + 256, j->in_off, j->n, j->gate); }
static void ws_order_scatter(void* a) { WsOrderJob* j = (WsOrderJob*)a; bro_order_scatter_kernel(j->in_off, j->n, j->scratch + 256, j->order); }
static void ws_sizes_finish(void* a) { WsOrderJob* j = (WsOrderJob*)a; bro_sizes_finish_kernel((int32_t*)j->order, j->n); }

extern "C" int bro_warpsim_order(const uint64_t* in_off, uint32_t n, uint32_t* order, uint32_t* gate2, int lane_order, uint64_t seed) {
    uint32_t scratch[512];
    memset(scratch, 0, sizeof(scratch));
    if (gate2) gate2[0] = gate2[1] = 0;
    WsOrderJob j; j.in_off = in_off; j.n = n; j.scratch = scratch; j.order = order; j.gate = gate2;
    WsWarp* w = (WsWarp*)malloc(sizeof(WsWarp));
    const uint32_t blocks = (n + 255u) / 256u;
    int err = 0;
    ws_tid_base = 0u;
    for (uint32_t b = 0; b < blocks && !err; b++) { ws_block_idx = b; err = ws_run(w, ws_order_hist, &j, lane_order, seed + b, 256); }
    ws_block_idx = 0;
    if (!err) err = ws_run(w, ws_order_scan, &j, lane_order, seed, 256);
    for (uint32_t b = 0; b < blocks && !err; b++) { ws_block_idx = b; err = ws_run(w, ws_order_scatter, &j, lane_order, seed + 7u * b, 256); }
    ws_block_idx = 0;
    free(w);
    return err;
}
// bro_sizes_finish_kernel over n statuses
extern "C" int bro_warpsim_sizes_finish(int32_t* status, uint32_t n) {
    WsOrderJob j; memset(&j, 0, sizeof(j)); j.order = (uint32_t*)status; j.n = n;
    WsWarp* w = (WsWarp*)malloc(sizeof(WsWarp));
    int err = 0;
    for (uint32_t b = 0; b < (n + 255u) / 256u && !err; b++) { ws_block_idx = b; err = ws_run(w, ws_sizes_finish, &j, 0, 1, 256); }
    ws_block_idx = 0;
    free(w);
    return err;
}

extern "C" uint64_t bro_warpsim_parse_last_rendezvous() { return g_parse_rendezvous; }
extern "C" unsigned bro_warpsim_parse_block_bytes() { return BRO_TL_BYTES; }
extern "C" void bro_warpsim_parse_ring_slack(unsigned groups) { g_ring_slack = groups; }

#if defined(BRO_WARPSIM_MAIN)
// warpsim_parse_tsan <lanes> <order 0|1|2> <seed> file:size...: the files are the compressed streams of ONE batch.  Prints
// "name status out_len records" per stream.  Under ThreadSanitizer the 32 lanes' accesses must not meet anywhere: every lane owns
// its stream, its output slot, its share of the record arena, its table arena and -- word by word between the others' -- its
// block of shared memory; a lane that reaches into a neighbour's block is a report.
int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s lanes order seed file:size...\n", argv[0]); return 2; }
    const uint32_t lanes = (uint32_t)atoi(argv[1]);
    const int order = atoi(argv[2]);
    const uint64_t seed = strtoull(argv[3], 0, 10);
    const uint32_t n = (uint32_t)(argc - 4);
    if (getenv("BRO_WS_SMEM_SKEW")) ws_smem_skew_lane = atoi(getenv("BRO_WS_SMEM_SKEW"));      // mutation: the report must come
    uint64_t* in_off = (uint64_t*)calloc(2u * (n + 1u), sizeof(uint64_t));
    uint64_t* out_off = in_off + (n + 1u);
    uint8_t* in = (uint8_t*)calloc(512, 1);
    const size_t PADB = 256;
    for (uint32_t i = 0; i < n; i++) {
        char* colon = strrchr(argv[4 + i], ':');
        if (!colon) return 2;
        *colon = 0;
        out_off[i + 1] = out_off[i] + strtoull(colon + 1, 0, 10);
        FILE* f = fopen(argv[4 + i], "rb");
        if (!f) { perror(argv[4 + i]); return 2; }
        fseek(f, 0, SEEK_END);
        const size_t len = (size_t)ftell(f);
        fseek(f, 0, SEEK_SET);
        in = (uint8_t*)realloc(in, PADB + (size_t)in_off[i] + len + PADB);
        if (fread(in + PADB + in_off[i], 1, len, f) != len) return 2;
        fclose(f);
        in_off[i + 1] = in_off[i] + len;
    }
    memset(in + PADB + in_off[n], 0xee, PADB);
    uint8_t* out = (uint8_t*)calloc((size_t)out_off[n] + 2 * PADB, 1);
    const uint64_t rec_total = BRO_REC_BASE(in_off, n);
    uint32_t* rec = (uint32_t*)aligned_alloc(16, 16u * (size_t)(rec_total + 1u));
    uint64_t* out_len = (uint64_t*)calloc(n + 1u, sizeof(uint64_t));
    int32_t* status = (int32_t*)calloc(n + 1u, sizeof(int32_t));
    uint32_t* nrec = (uint32_t*)calloc(n + 1u, sizeof(uint32_t));
    uint32_t* done_q = (uint32_t*)calloc(n + 1u, sizeof(uint32_t));
    uint32_t retry = 0;
    const int err = bro_warpsim_parse_launch(in + PADB, in_off, out + PADB, out_off, out_len, status, nrec, rec, rec_total, n, 0, lanes, 0, 0, order,
                                             seed, 1, done_q, &retry, getenv("BRO_WS_THREADS") ? atoi(getenv("BRO_WS_THREADS")) : 32);
    for (uint32_t i = 0; i < n; i++) printf("%s %d %llu %u %d\n", argv[4 + i], status[i], (unsigned long long)out_len[i], nrec[i], err);
    return err ? 3 : 0;
}
#endif
