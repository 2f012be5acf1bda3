// bro_warpsim_copy.cpp -- 32-LANE host simulation of the COPY KERNEL (phase two of the two-phase path) -- CPU TEST-SUITE ONLY.
// Built into tests/_build/libbro_warpsim_copy.so (and, with -fsanitize=thread -DBRO_WARPSIM_MAIN, into
// tests/_build/warpsim_copy_tsan) by tests/warpsim.py together with the host simulation of phase one
// (bro_hostsim_parse.cpp); never part of libbrotli_b200.so.
//
// bro_kernels_copy.cu is compiled HERE, unchanged, with g++ -- the kernel function itself: the completion queue, the
// split of 32 records into groups of independent records (ballot / ffs), the periodic fill by doubling, the long-record
// path (bro_run_pieces: lane groups of 8, all loads of four pieces before the first store) and the short-record path
// (prefix sum and binary search by shuffle, sources staged through shared memory by cp.async).  One warp = 32 fibers
// (bro_warpsim.h) runs a batch of streams whose records phase one has just written.  cp.async is modelled as the hardware
// allows it to behave: the 16 bytes move at the latest possible moment, the lane's cp.async.wait_all.
//
// What this adds to bro_hostsim_copy.cpp (a restatement of the kernel's control flow around its lane-local piece code):
// the warp-wide code itself, under three lane orders -- and, in the ThreadSanitizer build, a race check of it under the
// CUDA memory model in which only __syncwarp orders the lanes' accesses to the OUTPUT SLOT (global memory:
// compute-sanitizer's racecheck does not look there), which is what the kernel's correctness rests on.
#define BRO_WARPSIM 1
#define __CUDACC__ 1          /* bro_copy_piece.h: the device forms (uint4 vectors, 32-bit shared-window addresses) */
#include "bro_warpsim.h"

// the asynchronous copies of a lane (global -> shared, 16 bytes): issued now, performed at the lane's wait
struct WsAsyncCopy { uint32_t dst; const void* src; };
static WsAsyncCopy g_async[WS_MAX_THREADS][64];
static int g_async_n[WS_MAX_THREADS];
WS_NO_TSAN static int ws_async_push(uint32_t dst, const void* src) {
    const unsigned l = ws_tid();
    if (g_async_n[l] >= 64) abort();
    g_async[l][g_async_n[l]].dst = dst; g_async[l][g_async_n[l]].src = src;
    return g_async_n[l]++;
}
WS_NO_TSAN static int ws_async_take(WsAsyncCopy* out) {
    const unsigned l = ws_tid();
    const int n = g_async_n[l];
    memcpy(out, g_async[l], sizeof(WsAsyncCopy) * (size_t)n);
    g_async_n[l] = 0;
    return n;
}
static inline void bro_cp_async16(uint32_t smem_addr, const void* gptr) { (void)ws_async_push(smem_addr, gptr); }
static inline void bro_cp_async_wait_all() {
    WsAsyncCopy q[64];
    const int n = ws_async_take(q);
    for (int k = 0; k < n; k++) *(uint4*)ws_smem_ptr(q[k].dst) = *(const uint4*)q[k].src;      // (instrumented: this lane reads the source now)
}
static inline uint4 bro_lds128(uint32_t smem_addr) { return *(const uint4*)ws_smem_ptr(smem_addr); }
#define BRO_PREFETCH_BULK_L2(addr, bytes) ((void)(addr), (void)(bytes))
#define BRO_PREFETCH_L2(ptr) ((void)(ptr))

// the kernel's barriers by source line of bro_kernels_copy.cu (counted; one of them can be left out: a mutation the lane
// orders / the race detector must notice)
static inline void ws_syncwarp_real() { __syncwarp(0xffffffffu); }
static inline void ws_syncwarp_line(int line) {
    ws_sync_count(line);
    if (line == g_sync_drop_line) return;
    ws_syncwarp_real();
}
#define __syncwarp(...) ws_syncwarp_line(__LINE__)

#include "bro_kernels_copy.cu"
#undef __syncwarp

// ------------------------------------------------------------------------------------------------------
// one launch: a batch of n streams, one warp
// ------------------------------------------------------------------------------------------------------
struct WsCopyJob { BroLaunch p; int shape; };

static void ws_copy_lane(void* arg) {
    WsCopyJob* j = (WsCopyJob*)arg;
    if (j->shape) bro_copy_kernel<BRO_COPY_WARPS_SMALL, BRO_COPY_MIN_BLOCKS_SMALL>(j->p);
    else bro_copy_kernel<BRO_COPY_WARPS, BRO_COPY_MIN_BLOCKS>(j->p);
}

static uint64_t g_copy_rendezvous;

// Executes the records of n streams as the copy kernel does.  in / in_off, out / out_off, status, nrec, rec: as in BroLaunch
// (stream i owns the records from BRO_REC_BASE(in_off, i) on); the caller's `out` must be addressable 16 bytes either side of
// every slot the way the product's output allocation is.  queue_order: the completion queue holds the streams in this order
// (NULL = identity).  Returns the simulation's verdict (0 = the warp behaved); stats[0..1] = bytes moved, records executed.
extern "C" int bro_warpsim_copy_launch(const uint8_t* in, const uint64_t* in_off, uint8_t* out, const uint64_t* out_off, const int32_t* status,
                                       const uint32_t* nrec, const uint32_t* rec_words, uint32_t n, const uint32_t* queue_order, int shape,
                                       int order, uint64_t seed, unsigned long long* stats, int nthreads = 32) {
    WsCopyJob j;
    memset(&j, 0, sizeof(j));
    uint32_t counter = 0, fault = 0;
    uint32_t* done_q = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1u));
    for (uint32_t i = 0; i < n; i++) done_q[i] = queue_order ? queue_order[i] : i;
    unsigned long long cs[2] = {0, 0};
    j.p.in = in; j.p.in_off = in_off; j.p.out = out; j.p.out_off = out_off;
    j.p.status = (int32_t*)status; j.p.n = n; j.p.counter = &counter;
    j.p.rec = (BroRec*)rec_words; j.p.nrec = (uint32_t*)nrec; j.p.done_q = done_q;
    j.p.gate = 0; j.p.fault = &fault; j.p.watchdog = 1ll << 36; j.p.copy_stats = cs;
    j.shape = shape;
    memset(g_async_n, 0, sizeof(g_async_n));
    WsWarp* w = (WsWarp*)malloc(sizeof(WsWarp));
    if (nthreads > 32 * (shape ? BRO_COPY_WARPS_SMALL : BRO_COPY_WARPS)) abort();       // (the kernel's staging is sized for its block)
    int err = ws_run(w, ws_copy_lane, &j, order, seed, nthreads < 32 ? 32 : nthreads);
    g_copy_rendezvous = w->rendezvous;
    if (!err && fault) err = 101;
    if (stats) { stats[0] = cs[0]; stats[1] = cs[1]; }
    free(w); free(done_q);
    return err;
}
extern "C" uint64_t bro_warpsim_copy_last_rendezvous() { return g_copy_rendezvous; }

// ------------------------------------------------------------------------------------------------------
// the two-phase path for a batch: phase one = the parse kernel's per-lane code on the host (bro_hostsim_parse.cpp), stream by
// stream, writing literals and dictionary words into the slots and the copy records into the streams' shares of the record
// arena; phase two = one launch of the simulated copy kernel over the whole batch.
// ------------------------------------------------------------------------------------------------------
extern "C" void bro_hostsim_parse_set_copy_group(int group);
extern "C" void bro_hostsim_parse_export_records(uint32_t* words, unsigned max_records);
extern "C" int bro_hostsim_parse_decode(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap, size_t* out_len,
                                        int quirks, unsigned arena_u16, unsigned rec_cap, unsigned* n_rec, unsigned* n_steps);

// in / in_off / out_off as in bro_batch_decode (n + 1 offsets each); out: out_off[n] bytes, filled for the streams whose status is
// OK.  in_mis / out_mis: alignment (mod 16) of the first compressed byte and of the first slot in the simulation's own buffers.
// queue_seed != 0: the completion queue hands the streams out in a shuffled order.  Returns the simulation's verdict.
extern "C" int bro_warpsim_two_phase(const uint8_t* in, const uint64_t* in_off, uint8_t* out, const uint64_t* out_off, uint64_t* out_len,
                                     int32_t* status, uint32_t n, int quirks, int shape, int order, uint64_t seed, unsigned in_mis, unsigned out_mis,
                                     uint64_t queue_seed, unsigned long long* stats) {
    enum { PAD = 256 };
    const uint64_t in_bytes = in_off[n] - in_off[0], out_bytes = out_off[n] - out_off[0];
    uint8_t* inb = (uint8_t*)malloc(in_bytes + 2 * PAD + 16);
    uint8_t* outb = (uint8_t*)malloc(out_bytes + 2 * PAD + 16);
    memset(inb, 0xee, in_bytes + 2 * PAD + 16);
    memset(outb, 0xdd, out_bytes + 2 * PAD + 16);
    uint8_t* in_al = (uint8_t*)(((uintptr_t)inb + PAD + 15) & ~(uintptr_t)15) + (in_mis & 15u);
    uint8_t* out_al = (uint8_t*)(((uintptr_t)outb + PAD + 15) & ~(uintptr_t)15) + (out_mis & 15u);
    memcpy(in_al, in + in_off[0], in_bytes);
    // the launch sees offsets from the simulation's buffers
    uint64_t* ioff = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1u) * 2u);
    uint64_t* ooff = ioff + (n + 1u);
    for (uint32_t i = 0; i <= n; i++) { ioff[i] = in_off[i] - in_off[0]; ooff[i] = out_off[i] - out_off[0]; }
    const uint64_t rec_total = BRO_REC_BASE(ioff, n);
    uint32_t* rec = (uint32_t*)aligned_alloc(16, 16u * (size_t)(rec_total + 1u));
    memset(rec, 0xab, 16u * (size_t)(rec_total + 1u));
    uint32_t* nrec = (uint32_t*)calloc(n + 1u, sizeof(uint32_t));
    bro_hostsim_parse_set_copy_group(-1);
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t base = BRO_REC_BASE(ioff, i), share = BRO_REC_BASE(ioff, i + 1u) - base;
        size_t len = 0;
        unsigned nr = 0, steps = 0;
        bro_hostsim_parse_export_records(rec + 4u * base, (unsigned)share);
        status[i] = bro_hostsim_parse_decode(in_al + ioff[i], (size_t)(ioff[i + 1] - ioff[i]), out_al + ooff[i], (size_t)(ooff[i + 1] - ooff[i]), &len,
                                             quirks, 0u, (unsigned)share, &nr, &steps);
        out_len[i] = len;
        nrec[i] = nr;
        for (uint32_t k = 0; k < nr; k++) rec[4u * (base + k) + 3u] = 0u;      // (the export put the slot's alignment there; BroRec.b is 0)
    }
    bro_hostsim_parse_set_copy_group(0);
    uint32_t* queue = 0;
    if (queue_seed) {
        queue = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1u));
        for (uint32_t i = 0; i < n; i++) queue[i] = i;
        uint64_t r = queue_seed;
        for (uint32_t k = n; k > 1u; k--) {
            r = r * 6364136223846793005ull + 1442695040888963407ull;
            const uint32_t jx = (uint32_t)((r >> 33) % k);
            const uint32_t t = queue[k - 1u]; queue[k - 1u] = queue[jx]; queue[jx] = t;
        }
    }
    int err = bro_warpsim_copy_launch(in_al, ioff, out_al, ooff, status, nrec, rec, n, queue, shape, order, seed, stats);
    // nothing outside the slots may have been written
    bool clobber = false;
    for (uint8_t* p = outb; p < out_al && !clobber; p++) clobber = *p != 0xdd;
    for (uint8_t* p = out_al + out_bytes; p < outb + out_bytes + 2 * PAD + 16 && !clobber; p++) clobber = *p != 0xdd;
    if (clobber && !err) err = 100;
    memcpy(out + out_off[0], out_al, out_bytes);
    free(queue); free(nrec); free(rec); free(ioff); free(outb); free(inb);
    return err;
}

extern "C" void bro_warpsim_copy_sync_hits(uint64_t* hits, int n, int reset) {
    for (int i = 0; i < n && i < WS_MAX_LINE; i++) hits[i] = g_sync_hits[i];
    if (reset) memset(g_sync_hits, 0, sizeof(g_sync_hits));
}
extern "C" void bro_warpsim_copy_drop_sync(int line) { g_sync_drop_line = line; }

#if defined(BRO_WARPSIM_MAIN)
// warpsim_copy_tsan <shape 0|1> <order 0|1|2> <seed> <queue_seed> file:size...: the files are the compressed streams of ONE batch,
// each decoded into a slot of the size given.  Prints "name status out_len fnv1a64(out)" per stream; ThreadSanitizer's reports go
// to stderr and make the exit code non-zero.  BRO_WS_ALIGN=in,out  BRO_WS_DROP_SYNC=line as in warpsim_tsan.
int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s shape order seed queue_seed file:size...\n", argv[0]); return 2; }
    const int shape = atoi(argv[1]), order = atoi(argv[2]);
    const uint64_t seed = strtoull(argv[3], 0, 10), queue_seed = strtoull(argv[4], 0, 10);
    unsigned in_mis = 0, out_mis = 0;
    if (getenv("BRO_WS_ALIGN")) sscanf(getenv("BRO_WS_ALIGN"), "%u,%u", &in_mis, &out_mis);
    if (getenv("BRO_WS_DROP_SYNC")) g_sync_drop_line = atoi(getenv("BRO_WS_DROP_SYNC"));
    const uint32_t n = (uint32_t)(argc - 5);
    uint64_t* in_off = (uint64_t*)calloc(2u * (n + 1u), sizeof(uint64_t));
    uint64_t* out_off = in_off + (n + 1u);
    uint8_t* in = 0;
    for (uint32_t i = 0; i < n; i++) {
        char* colon = strrchr(argv[5 + i], ':');
        if (!colon) return 2;
        *colon = 0;
        out_off[i + 1] = out_off[i] + strtoull(colon + 1, 0, 10);
        FILE* f = fopen(argv[5 + i], "rb");
        if (!f) { perror(argv[5 + i]); return 2; }
        fseek(f, 0, SEEK_END);
        const size_t len = (size_t)ftell(f);
        fseek(f, 0, SEEK_SET);
        in = (uint8_t*)realloc(in, (size_t)in_off[i] + len + 1);
        if (fread(in + in_off[i], 1, len, f) != len) return 2;
        fclose(f);
        in_off[i + 1] = in_off[i] + len;
    }
    uint8_t* out = (uint8_t*)calloc((size_t)out_off[n] + 1, 1);
    uint64_t* out_len = (uint64_t*)calloc(n, sizeof(uint64_t));
    int32_t* status = (int32_t*)calloc(n, sizeof(int32_t));
    unsigned long long stats[2];
    const int err = bro_warpsim_two_phase(in, in_off, out, out_off, out_len, status, n, 0, shape, order, seed, in_mis, out_mis, queue_seed, stats);
    for (uint32_t i = 0; i < n; i++) {
        uint64_t h = 1469598103934665603ull;
        const uint64_t cap = out_off[i + 1] - out_off[i], len = out_len[i] < cap ? out_len[i] : cap;     // (a failed stream may report a position beyond its slot)
        for (uint64_t k = 0; k < len; k++) h = (h ^ out[out_off[i] + k]) * 1099511628211ull;
        printf("%s %d %llu %016llx %d\n", argv[5 + i], status[i], (unsigned long long)out_len[i], (unsigned long long)h, err);
    }
    return err ? 3 : 0;
}
#endif
