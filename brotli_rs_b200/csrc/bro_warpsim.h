// bro_warpsim.h -- a CTA of one to eight warps on the host -- CPU TEST-SUITE ONLY (bro_warpsim.cpp: the fused and resume kernels;
// bro_warpsim_parse.cpp: the parse and ordering kernels; bro_warpsim_copy.cpp: the copy kernel).  Never part of libbrotli_b200.so.
//
// The threads are fibers (one stack each) on one OS thread; the warp intrinsics are rendezvous points between the lanes of a
// warp, __syncthreads between all threads.  Between two rendezvous a thread runs alone, in an order the caller chooses (ascending,
// descending, a seeded shuffle re-drawn at every rendezvous; the warps of a CTA are interleaved thread by thread).  Checked
// along the way: every lane arrives at the SAME intrinsic with the same mask, no thread leaves while others wait for it, all
// threads return the same result.  Built with -fsanitize=thread the threads are ThreadSanitizer fibers and only __syncwarp /
// __syncthreads order memory between them (see below).  Also in here: stand-ins for the rest of the CUDA surface the kernels
// use (vector types, atomics, 32-bit shared-window addresses), and buffers that end at unmapped pages.
#pragma once
#if !defined(__x86_64__)
#error "bro_warpsim.h: the fiber switch below is x86-64 only (tests/warpsim.py skips the suite elsewhere)"
#endif
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

enum { WS_LANES = 32, WS_MAX_THREADS = 256, WS_STACK = 512 * 1024 };     // a CTA of up to 8 warps
enum { WS_OP_NONE = 0, WS_OP_SHFL, WS_OP_MATCH, WS_OP_ALL, WS_OP_ANY, WS_OP_BALLOT, WS_OP_SYNC, WS_OP_SHFL_UP, WS_OP_SHFL_XOR, WS_OP_ADD, WS_OP_MAX, WS_OP_SYNC_CTA };
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
struct WsWarp {                          // (a CTA: nthreads / 32 warps; the name is from when it was one)
    WsLane lane[WS_MAX_THREADS];
    int nthreads;
    int cur;
    void* main_sp;
    int order_mode;          // 0 ascending, 1 descending, 2 shuffled at every rendezvous
    uint64_t rng;
    int perm[WS_MAX_THREADS];      // perm[k] = the thread that runs k-th
    int where[WS_MAX_THREADS];     // inverse
    int err;
    uint64_t rendezvous;
    void (*body)(void*);
    void* arg;
    void* main_fiber;
    char sync_token[WS_MAX_THREADS / 32], cta_token, start_token, end_token;     // addresses the happens-before edges hang on
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
    const int N = w->nthreads;
    for (int k = 0; k < N; k++) w->perm[k] = w->order_mode == 1 ? N - 1 - k : k;
    if (w->order_mode == 2)
        for (int k = N - 1; k > 0; k--) {
            w->rng = w->rng * 6364136223846793005ull + 1442695040888963407ull;
            const int j = (int)((w->rng >> 33) % (uint64_t)(k + 1));
            const int t = w->perm[k]; w->perm[k] = w->perm[j]; w->perm[j] = t;
        }
    for (int k = 0; k < N; k++) w->where[w->perm[k]] = k;
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
    for (int n = 0; n < w->nthreads; n++) {
        k = (k + 1) % w->nthreads;
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
    // the threads this one meets: the lanes of its warp that the mask names -- or, at __syncthreads, the whole CTA
    const bool cta = op == WS_OP_SYNC_CTA;
    WsLane* const wl = cta ? w->lane : &w->lane[me->tid & ~31u];       // lane 0 of the set
    const int cnt = cta ? w->nthreads : WS_LANES;
    const unsigned lane = cta ? me->tid : (me->tid & 31u);
    if (!cta && !((mask >> lane) & 1u)) ws_fail(w, WS_ERR_BAD_MASK);
#define WS_IN(l) (cta || ((mask >> (l)) & 1u))
    me->waiting = 1; me->op = op; me->mask = mask; me->a = a; me->b = b;
    for (;;) {
        // have all the threads this one names arrived?
        bool all = true;
        for (int l = 0; l < cnt && all; l++)
            if (WS_IN(l)) {
                if (wl[l].done) ws_fail(w, WS_ERR_EXIT_WHILE_WAITED);
                if (!wl[l].waiting || (wl[l].op == WS_OP_SYNC_CTA) != cta) all = false;
            }
        if (all) break;
        ws_yield(w);
        if (!me->waiting) return me->result;      // the last arriver released this thread
        // a full round without progress cannot happen silently: a thread that runs either arrives, returns or fails
    }
    // this thread is the last to arrive: same intrinsic, same mask everywhere, then compute and release
    for (int l = 0; l < cnt; l++)
        if (WS_IN(l) && (wl[l].op != op || wl[l].mask != mask)) ws_fail(w, WS_ERR_DIVERGENT);
    uint32_t ballot = 0;
    if (!cta) for (int l = 0; l < WS_LANES; l++) if (((mask >> l) & 1u) && wl[l].a) ballot |= 1u << l;
    for (int l = 0; l < cnt; l++) {
        if (!WS_IN(l)) continue;
        WsLane* t = &wl[l];
        uint32_t r = 0;
        switch (op) {
        case WS_OP_SHFL: {
            const uint32_t width = t->b >> 8, src = t->b & 255u;
            const uint32_t from = ((uint32_t)l & ~(width - 1u)) | (src & (width - 1u));
            r = ((mask >> from) & 1u) ? wl[from].a : t->a;     // (a lane outside the mask: undefined on the device)
            break;
        }
        case WS_OP_SHFL_UP: {
            const uint32_t width = t->b >> 8, delta = t->b & 255u;
            const uint32_t in_seg = (uint32_t)l & (width - 1u);
            r = in_seg >= delta && ((mask >> (l - (int)delta)) & 1u) ? wl[l - (int)delta].a : t->a;
            break;
        }
        case WS_OP_SHFL_XOR: {
            const uint32_t from = (uint32_t)l ^ (t->b & 255u);
            r = from < WS_LANES && ((mask >> from) & 1u) ? wl[from].a : t->a;
            break;
        }
        case WS_OP_ADD:
            for (int j = 0; j < WS_LANES; j++) if ((mask >> j) & 1u) r += wl[j].a;
            break;
        case WS_OP_MAX:
            for (int j = 0; j < WS_LANES; j++) if (((mask >> j) & 1u) && wl[j].a > r) r = wl[j].a;
            break;
        case WS_OP_MATCH:
            for (int j = 0; j < WS_LANES; j++) if (((mask >> j) & 1u) && wl[j].a == t->a) r |= 1u << j;
            break;
        case WS_OP_ALL: r = ballot == mask; break;
        case WS_OP_ANY: r = ballot != 0u; break;
        case WS_OP_BALLOT: r = ballot; break;
        default: break;
        }
        t->result = r;
        t->waiting = 0;
    }
#undef WS_IN
    w->rendezvous++;
    if (w->order_mode == 2) ws_set_order(w);
    return me->result;
}

// __syncwarp (and __syncthreads for the CTA) are the intrinsics that order memory: everything a thread did before one
// happens-before everything any thread it met there does after it
WS_NO_TSAN static uint32_t ws_collective(uint32_t op, uint32_t mask, uint32_t a, uint32_t b) {
#if WS_TSAN
    char* const token = op == WS_OP_SYNC_CTA ? &g_ws->cta_token : &g_ws->sync_token[g_ws->lane[g_ws->cur].tid >> 5];
    if (op == WS_OP_SYNC || op == WS_OP_SYNC_CTA) __tsan_release(token);
#endif
    const uint32_t r = ws_collective_raw(op, mask, a, b);
#if WS_TSAN
    if (op == WS_OP_SYNC || op == WS_OP_SYNC_CTA) __tsan_acquire(token);
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
    // a thread must not leave while another waits for it at a rendezvous (a lane of its warp that names it, or anyone at __syncthreads)
    for (int l = 0; l < w->nthreads; l++) {
        const WsLane* o = &w->lane[l];
        if (!o->waiting) continue;
        if (o->op == WS_OP_SYNC_CTA || ((o->tid >> 5) == (me->tid >> 5) && ((o->mask >> (me->tid & 31u)) & 1u))) ws_fail(w, WS_ERR_EXIT_WHILE_WAITED);
    }
    ws_yield(w);                // to the next live lane, or to main when this was the last
    abort();
}

// nthreads: a multiple of 32 up to WS_MAX_THREADS (the CTA's warps run interleaved, thread by thread, in the chosen order)
WS_NO_TSAN static int ws_run(WsWarp* w, void (*body)(void*), void* arg, int order_mode, uint64_t seed, int nthreads = WS_LANES) {
    memset(w, 0, sizeof(*w));
    if (nthreads < WS_LANES || nthreads > WS_MAX_THREADS || (nthreads & 31)) abort();
    w->nthreads = nthreads;
    w->body = body; w->arg = arg; w->order_mode = order_mode; w->rng = seed * 2654435761ull + 1ull;
    ws_set_order(w);
    for (int l = 0; l < nthreads; l++) {
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
    for (int l = 0; l < nthreads; l++) __tsan_destroy_fiber(w->lane[l].fiber);
#endif
    g_ws = 0;
    int alive = 0;
    for (int l = 0; l < nthreads; l++) { alive += !w->lane[l].done; free(w->lane[l].stack); }
    if (!w->err && alive) w->err = WS_ERR_DIVERGENT;
    if (!w->err)
        for (int l = 1; l < nthreads; l++)
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
struct alignas(16) uint4 { uint32_t x, y, z, w; };     // a misaligned vector access faults here as it does on the device
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = {x, y, z, w}; return v; }
struct WsTid { unsigned x; };
WS_NO_TSAN static unsigned ws_tid() { return g_ws->lane[g_ws->cur].tid; }
#define threadIdx (WsTid{ws_tid()})
static unsigned ws_block_idx;         // the CTA being run (a grid is run CTA after CTA)
#define blockIdx (WsTid{ws_block_idx})
#define blockDim (WsTid{(unsigned)g_ws->nthreads})
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
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline uint32_t __shfl_up_sync(uint32_t mask, uint32_t v, unsigned delta, int width = 32) {
    return ws_collective(WS_OP_SHFL_UP, mask, v, ((uint32_t)width << 8) | (delta & 255u));
}
static inline uint32_t __shfl_xor_sync(uint32_t mask, uint32_t v, int lanemask, int width = 32) {
    return ws_collective(WS_OP_SHFL_XOR, mask, v, ((uint32_t)width << 8) | ((uint32_t)lanemask & 255u));
}
// 64-bit values travel as two 32-bit halves (as the hardware moves them)
static inline uint64_t __shfl_sync(uint32_t mask, uint64_t v, int src, int width = 32) {
    const uint32_t lo = __shfl_sync(mask, (uint32_t)v, src, width), hi = __shfl_sync(mask, (uint32_t)(v >> 32), src, width);
    return lo | ((uint64_t)hi << 32);
}
static inline unsigned long long __shfl_xor_sync(uint32_t mask, unsigned long long v, int lanemask, int width = 32) {
    const uint32_t lo = __shfl_xor_sync(mask, (uint32_t)v, lanemask, width), hi = __shfl_xor_sync(mask, (uint32_t)(v >> 32), lanemask, width);
    return lo | ((unsigned long long)hi << 32);
}
static inline uint32_t __reduce_add_sync(uint32_t mask, uint32_t v) { return ws_collective(WS_OP_ADD, mask, v, 0); }
static inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) { return ws_collective(WS_OP_MAX, mask, v, 0); }
static inline void __syncthreads() { (void)ws_collective(WS_OP_SYNC_CTA, 0, 0, 0); }
WS_NO_TSAN static uint32_t atomicMax(uint32_t* p, uint32_t v) { const uint32_t o = *p; if (v > o) *p = v; return o; }
// atomics and the rest of the kernel-level surface: one warp runs at a time, so a plain read-modify-write is atomic here
// (not instrumented: on the device these are atomic operations, not data accesses)
WS_NO_TSAN static uint32_t atomicAdd(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
WS_NO_TSAN static unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
WS_NO_TSAN static uint32_t atomicExch(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = v; return o; }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) {}
static inline void __threadfence() {}
#define __global__ static
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
typedef void* cudaStream_t;
// shared memory is addressed through 32-bit window addresses in the kernels' PTX: here an offset from a fixed anchor
static uint8_t ws_smem_anchor[16] __attribute__((aligned(16)));
static inline uintptr_t ws_smem_base() { return (uintptr_t)ws_smem_anchor - 0x40000000u; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)(uint32_t)((uintptr_t)p - ws_smem_base()); }
static int ws_smem_skew_lane = -1;      // mutation: this lane's shared-memory accesses land 4 bytes further on -- in its neighbour's words
static inline uint8_t* ws_smem_ptr(uint32_t a) { return (uint8_t*)(ws_smem_base() + a) + (ws_smem_skew_lane >= 0 && (int)ws_tid() == ws_smem_skew_lane ? 4 : 0); }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------------
// Guarded buffers: how far beyond what it was given does a kernel READ?  (ws_guard_mode: 0 none; 1 = the byte `slack` bytes
// behind the buffer's end is the first byte of an unmapped page; 2 = the buffer starts a page whose predecessor is unmapped.)
// A touch of the unmapped page ends the process with exit code 97 (the drivers with a main() install the handler).
// ------------------------------------------------------------------------------------------------------
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>
static int ws_guard_mode;
static inline uint8_t* ws_guard_alloc(size_t bytes, size_t slack, unsigned mis, uint8_t fill) {
    const size_t page = 4096;
    const size_t body = (bytes + slack + mis + 2 * page - 1) / page * page;
    uint8_t* m = (uint8_t*)mmap(0, body + 2 * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == (uint8_t*)MAP_FAILED) abort();
    memset(m, fill, body + 2 * page);
    mprotect(m, page, PROT_NONE);
    mprotect(m + page + body, page, PROT_NONE);
    return ws_guard_mode == 2 ? m + page + mis : m + page + body - slack - bytes;      // (never unmapped: test processes are short-lived)
}
static void ws_guard_handler(int, siginfo_t* si, void*) {
    char msg[96];
    const int n = snprintf(msg, sizeof(msg), "GUARD: access at %p outside the buffers a caller must provide\n", si->si_addr);
    if (write(2, msg, (size_t)n) < 0) {}
    _exit(97);
}
static inline void ws_guard_install() {
    static uint8_t alt[65536];
    stack_t ss; ss.ss_sp = alt; ss.ss_size = sizeof(alt); ss.ss_flags = 0;
    sigaltstack(&ss, 0);
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_sigaction = ws_guard_handler; sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
    sigaction(SIGSEGV, &sa, 0);
}

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
