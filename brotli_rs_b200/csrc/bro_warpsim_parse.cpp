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
static uint8_t ws_dynamic_smem[2 * 32 * 1024] __attribute__((aligned(128)));
static unsigned ws_tid_base;
struct WsTidB { unsigned x; };
#undef threadIdx
#define threadIdx (WsTidB{ws_tid() + ws_tid_base})
static inline void __syncthreads() { __syncwarp(0xffffffffu); }      // (one warp of the CTA runs at a time)

// the ring's asynchronous copies (global -> shared, 4 bytes, one commit group each), per lane, oldest first
struct WsRingCopy { uint32_t dst; const void* src; };
static WsRingCopy g_ring[WS_LANES][64];
static unsigned g_ring_head[WS_LANES], g_ring_n[WS_LANES];
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
                                        uint32_t* retry_count) {
    BroLaunch p;
    memset(&p, 0, sizeof(p));
    uint32_t counter = 0, done_tail = 0, retry = 0;
    uint16_t* arena = (uint16_t*)aligned_alloc(128, (2u * (size_t)BRO_THREAD_ARENA_STRIDE_U16 * 64u + 127u) & ~(size_t)127);
    p.in = in; p.in_off = in_off; p.out = out; p.out_off = out_off; p.out_len = out_len; p.status = status;
    p.arena = arena; p.dict = bro_dictionary_blob; p.counter = &counter; p.order = order; p.retry_count = &retry;
    p.quirk_spec = quirks; p.rec = (BroRec*)rec_words; p.rec_total = rec_total; p.nrec = nrec; p.done_q = done_q; p.done_tail = &done_tail;
    p.gate = 0; p.sizing = sizing; p.lanes = lanes;
    for (uint32_t i = 0; i < n; i++) { status[i] = -12345; if (done_q) done_q[i] = 0xffffffffu; }
    g_ring_late = ring_late;
    memset(g_ring_n, 0, sizeof(g_ring_n));
    memset(ws_dynamic_smem, 0xcc, sizeof(ws_dynamic_smem));
    WsWarp* w = (WsWarp*)malloc(sizeof(WsWarp));
    // the CTA's second warp: it finds no stream (n = 0 for it) and leaves after its share of the insert/copy table
    p.n = 0;
    ws_tid_base = 32u;
    int err = ws_run(w, ws_parse_lane, &p, 0, 1);
    // the warp that decodes
    p.n = n;
    counter = 0;
    ws_tid_base = 0u;
    memset(g_ring_n, 0, sizeof(g_ring_n));
    if (!err) err = ws_run(w, ws_parse_lane, &p, lane_order, seed);
    g_parse_rendezvous = w->rendezvous;
    // every stream reported exactly once, and announced exactly once
    if (!err) {
        for (uint32_t i = 0; i < n && !err; i++) if (status[i] == -12345) err = 102;
        if (done_q && done_tail != n) err = 103;
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
extern "C" uint64_t bro_warpsim_parse_last_rendezvous() { return g_parse_rendezvous; }
extern "C" unsigned bro_warpsim_parse_block_bytes() { return BRO_TL_BYTES; }
extern "C" void bro_warpsim_parse_ring_slack(unsigned groups) { g_ring_slack = groups; }
