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

// ------------------------------------------------------------------------------------------------------
// Race detection (tests/_build/warpsim_tsan, built with -fsanitize=thread -DBRO_WARPSIM_MAIN): every lane is a
// ThreadSanitizer fiber, fiber switches carry NO happens-before edge, and the only edge between lanes is __syncwarp --
// which is the CUDA memory model for a warp: shuffles, votes and match move register values and order nothing.  A byte
// one lane stores and another loads (or stores) without a __syncwarp between the two is reported with both source
// lines.  compute-sanitizer's racecheck sees shared memory only; here the output slot and the table arena (global
// memory on the device) are covered as well.  The simulation's own bookkeeping is not instrumented.
// ------------------------------------------------------------------------------------------------------
#if defined(__SANITIZE_THREAD__)
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
void __tsan_acquire(void* addr);
void __tsan_release(void* addr);
}
#define WS_NO_TSAN __attribute__((no_sanitize("thread"), noinline))
#define WS_TSAN 1
#else
#define WS_NO_TSAN
#define WS_TSAN 0
#endif

// ------------------------------------------------------------------------------------------------------
// fibers
// ------------------------------------------------------------------------------------------------------
extern "C" void bro_ws_switch(void** save_sp, void* load_sp);
asm(".text\n"
    ".globl bro_ws_switch\n"
    ".type bro_ws_switch,@function\n"
    "bro_ws_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size bro_ws_switch,.-bro_ws_switch\n");

enum { WS_LANES = 32, WS_STACK = 512 * 1024 };
enum { WS_OP_NONE = 0, WS_OP_SHFL, WS_OP_MATCH, WS_OP_ALL, WS_OP_ANY, WS_OP_BALLOT, WS_OP_SYNC };
enum { WS_ERR_NONE = 0, WS_ERR_DIVERGENT = 1, WS_ERR_EXIT_WHILE_WAITED = 2, WS_ERR_NOT_UNIFORM = 3, WS_ERR_BAD_MASK = 4 };

struct WsLane {
    void* sp;
    uint8_t* stack;
    unsigned tid;
    int done, waiting;
    uint32_t op, mask, a, b, result;
    int ret_status;
    uint32_t ret_pos;
    void* fiber;             // ThreadSanitizer's view of this lane (race-detection build)
};
struct WsWarp {
    WsLane lane[WS_LANES];
    int cur;
    void* main_sp;
    int order_mode;          // 0 ascending, 1 descending, 2 shuffled at every rendezvous
    uint64_t rng;
    int perm[WS_LANES];      // perm[k] = the lane that runs k-th
    int where[WS_LANES];     // inverse
    int err;
    uint64_t rendezvous;
    void (*body)(void*);
    void* arg;
    void* main_fiber;
    char sync_token, start_token, end_token;     // addresses the happens-before edges hang on
};
static WsWarp* g_ws;

// switch stacks: `to` < 0 = the caller of ws_run
WS_NO_TSAN static void ws_switch_to(WsWarp* w, void** save_sp, int to) {
#if WS_TSAN
    __tsan_switch_to_fiber(to < 0 ? w->main_fiber : w->lane[to].fiber, 1u /* no synchronisation */);
#endif
    bro_ws_switch(save_sp, to < 0 ? w->main_sp : w->lane[to].sp);
}

WS_NO_TSAN static void ws_set_order(WsWarp* w) {
    for (int k = 0; k < WS_LANES; k++) w->perm[k] = w->order_mode == 1 ? WS_LANES - 1 - k : k;
    if (w->order_mode == 2)
        for (int k = WS_LANES - 1; k > 0; k--) {
            w->rng = w->rng * 6364136223846793005ull + 1442695040888963407ull;
            const int j = (int)((w->rng >> 33) % (uint64_t)(k + 1));
            const int t = w->perm[k]; w->perm[k] = w->perm[j]; w->perm[j] = t;
        }
    for (int k = 0; k < WS_LANES; k++) w->where[w->perm[k]] = k;
}

// the run is over (clean or not): back to the caller of ws_run
WS_NO_TSAN static void ws_to_main(WsWarp* w) {
    WsLane* me = &w->lane[w->cur];
    ws_switch_to(w, &me->sp, -1);
}

// hand the processor to the next lane (in the current order) that has not returned
WS_NO_TSAN static void ws_yield(WsWarp* w) {
    const int from = w->cur;
    int k = w->where[from];
    for (int n = 0; n < WS_LANES; n++) {
        k = (k + 1) % WS_LANES;
        const int l = w->perm[k];
        if (!w->lane[l].done) {
            if (l == from) return;
            w->cur = l;
            ws_switch_to(w, &w->lane[from].sp, l);
            return;
        }
    }
    ws_to_main(w);
}

WS_NO_TSAN static void ws_fail(WsWarp* w, int err) {
    if (!w->err) w->err = err;
    ws_to_main(w);              // never resumed
    abort();
}

WS_NO_TSAN static uint32_t ws_collective_raw(uint32_t op, uint32_t mask, uint32_t a, uint32_t b) {
    WsWarp* w = g_ws;
    WsLane* me = &w->lane[w->cur];
    if (!((mask >> me->tid) & 1u)) ws_fail(w, WS_ERR_BAD_MASK);
    me->waiting = 1; me->op = op; me->mask = mask; me->a = a; me->b = b;
    for (;;) {
        // have all the lanes this one names arrived?
        bool all = true;
        for (int l = 0; l < WS_LANES && all; l++)
            if ((mask >> l) & 1u) {
                if (w->lane[l].done) ws_fail(w, WS_ERR_EXIT_WHILE_WAITED);
                if (!w->lane[l].waiting) all = false;
            }
        if (all) break;
        ws_yield(w);
        if (!me->waiting) return me->result;      // the last arriver released this lane
        // a full round without progress cannot happen silently: a lane that runs either arrives, returns or fails
    }
    // this lane is the last to arrive: same intrinsic, same mask everywhere, then compute and release
    for (int l = 0; l < WS_LANES; l++)
        if (((mask >> l) & 1u) && (w->lane[l].op != op || w->lane[l].mask != mask)) ws_fail(w, WS_ERR_DIVERGENT);
    uint32_t ballot = 0;
    for (int l = 0; l < WS_LANES; l++) if (((mask >> l) & 1u) && w->lane[l].a) ballot |= 1u << l;
    for (int l = 0; l < WS_LANES; l++) {
        if (!((mask >> l) & 1u)) continue;
        WsLane* t = &w->lane[l];
        uint32_t r = 0;
        switch (op) {
        case WS_OP_SHFL: {
            const uint32_t width = t->b >> 8, src = t->b & 255u;
            const uint32_t from = ((uint32_t)l & ~(width - 1u)) | (src & (width - 1u));
            r = ((mask >> from) & 1u) ? w->lane[from].a : t->a;     // (a lane outside the mask: undefined on the device)
            break;
        }
        case WS_OP_MATCH:
            for (int j = 0; j < WS_LANES; j++) if (((mask >> j) & 1u) && w->lane[j].a == t->a) r |= 1u << j;
            break;
        case WS_OP_ALL: r = ballot == mask; break;
        case WS_OP_ANY: r = ballot != 0u; break;
        case WS_OP_BALLOT: r = ballot; break;
        default: break;
        }
        t->result = r;
        t->waiting = 0;
    }
    w->rendezvous++;
    if (w->order_mode == 2) ws_set_order(w);
    return me->result;
}

// __syncwarp is the one intrinsic that orders memory: everything a lane did before it happens-before everything any lane
// does after it
WS_NO_TSAN static uint32_t ws_collective(uint32_t op, uint32_t mask, uint32_t a, uint32_t b) {
#if WS_TSAN
    if (op == WS_OP_SYNC) __tsan_release(&g_ws->sync_token);
#endif
    const uint32_t r = ws_collective_raw(op, mask, a, b);
#if WS_TSAN
    if (op == WS_OP_SYNC) __tsan_acquire(&g_ws->sync_token);
#endif
    return r;
}

WS_NO_TSAN static void ws_trampoline() {
    WsWarp* w = g_ws;
#if WS_TSAN
    __tsan_acquire(&w->start_token);           // what the caller prepared (buffers, the job) is visible to every lane
#endif
    w->body(w->arg);
#if WS_TSAN
    __tsan_release(&w->end_token);
#endif
    WsLane* me = &w->lane[w->cur];
    me->done = 1;
    // a lane must not leave while another waits for it at a rendezvous
    for (int l = 0; l < WS_LANES; l++)
        if (w->lane[l].waiting && ((w->lane[l].mask >> me->tid) & 1u)) ws_fail(w, WS_ERR_EXIT_WHILE_WAITED);
    ws_yield(w);                // to the next live lane, or to main when this was the last
    abort();
}

WS_NO_TSAN static int ws_run(WsWarp* w, void (*body)(void*), void* arg, int order_mode, uint64_t seed) {
    memset(w, 0, sizeof(*w));
    w->body = body; w->arg = arg; w->order_mode = order_mode; w->rng = seed * 2654435761ull + 1ull;
    ws_set_order(w);
    for (int l = 0; l < WS_LANES; l++) {
        WsLane* t = &w->lane[l];
        t->tid = (unsigned)l;
        t->stack = (uint8_t*)malloc(WS_STACK);
        uintptr_t top = ((uintptr_t)t->stack + WS_STACK) & ~(uintptr_t)15;
        uint64_t* s = (uint64_t*)(top - 64);       // six callee-saved registers, the entry address, one slot of padding
        memset(s, 0, 64);
        s[6] = (uint64_t)(uintptr_t)&ws_trampoline;
        t->sp = s;
#if WS_TSAN
        t->fiber = __tsan_create_fiber(0);
#endif
    }
    g_ws = w;
    w->cur = w->perm[0];
#if WS_TSAN
    w->main_fiber = __tsan_get_current_fiber();
    __tsan_release(&w->start_token);
#endif
    ws_switch_to(w, &w->main_sp, w->cur);
#if WS_TSAN
    __tsan_acquire(&w->end_token);
    for (int l = 0; l < WS_LANES; l++) __tsan_destroy_fiber(w->lane[l].fiber);
#endif
    g_ws = 0;
    int alive = 0;
    for (int l = 0; l < WS_LANES; l++) { alive += !w->lane[l].done; free(w->lane[l].stack); }
    if (!w->err && alive) w->err = WS_ERR_DIVERGENT;
    if (!w->err)
        for (int l = 1; l < WS_LANES; l++)
            if (w->lane[l].ret_status != w->lane[0].ret_status || w->lane[l].ret_pos != w->lane[0].ret_pos) w->err = WS_ERR_NOT_UNIFORM;
    return w->err;
}

// ------------------------------------------------------------------------------------------------------
// the CUDA surface bro_decoder_core.h uses in its 32-lane form
// ------------------------------------------------------------------------------------------------------
#define __device__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = {x, y, z, w}; return v; }
struct WsTid { unsigned x; };
WS_NO_TSAN static unsigned ws_tid() { return g_ws->lane[g_ws->cur].tid; }
#define threadIdx (WsTid{ws_tid()})
static inline uint32_t __shfl_sync(uint32_t mask, uint32_t v, int src, int width = 32) {
    return ws_collective(WS_OP_SHFL, mask, v, ((uint32_t)width << 8) | ((uint32_t)src & 255u));
}
static inline uint32_t __match_any_sync(uint32_t mask, uint32_t v) { return ws_collective(WS_OP_MATCH, mask, v, 0); }
static inline int __all_sync(uint32_t mask, int p) { return (int)ws_collective(WS_OP_ALL, mask, p != 0, 0); }
static inline int __any_sync(uint32_t mask, int p) { return (int)ws_collective(WS_OP_ANY, mask, p != 0, 0); }
static inline uint32_t __ballot_sync(uint32_t mask, int p) { return ws_collective(WS_OP_BALLOT, mask, p != 0, 0); }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { (void)ws_collective(WS_OP_SYNC, mask, 0, 0); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline uint32_t __brev(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {     // shift taken modulo 32, as the device does
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
}
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }

// bro_syncwarp() by source line: how often each barrier of the header ran, and one line whose barrier is left out (a
// mutation: the test-suite checks that the lane orders notice -- tests/test_warpsim_parity.py)
enum { WS_MAX_LINE = 4096 };
static uint64_t g_sync_hits[WS_MAX_LINE];
static int g_sync_drop_line = -1;
static void ws_sync_count(int line);
static inline void bro_ws_syncwarp_at(int line) {
    ws_sync_count(line);
    if (line == g_sync_drop_line) return;
    __syncwarp(0xffffffffu);
}

WS_NO_TSAN static void ws_sync_count(int line) { if (line >= 0 && line < WS_MAX_LINE) g_sync_hits[line]++; }

#include "bro_decoder_core.h"

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
    BroResume* ck;            // != 0: the resume kernel's body
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
        st = BRO_ST_OutputTooSmall;
        d.pos = j->ck->pos;
        if (d.pos <= d.cap) st = bro_decode_stream_resume(d, j->ck, j->in, j->in + j->in_len);
        __syncwarp();
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
    if (ck) memcpy(out_al, out, ck->pos <= cap ? ck->pos : cap);         // the history
    WsJob j;
    j.in = in_al; j.in_len = in_len; j.out = out_al; j.cap = cap; j.quirks = quirks; j.latency = latency; j.ck = ck;
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
    for (uint8_t* p = outb; p < out_al && !clobber; p++) clobber = *p != 0xdd;
    for (uint8_t* p = out_al + cap; p < outb + cap + 2 * PAD + 16 && !clobber; p++) clobber = *p != 0xdd;
    if (clobber && !*sim_err) *sim_err = 100;
    if (n > cap) n = cap;
    memcpy(out, out_al, ck ? (w->lane[0].ret_pos <= cap && !err ? cap : 0) : n);
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
    if (getenv("BRO_WS_DROP_SYNC")) g_sync_drop_line = atoi(getenv("BRO_WS_DROP_SYNC"));      // mutation: the report must name it
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
