// bro_abi.cu -- the extern "C" entry points declared in include/brotli_b200.h.
//
// Host side of the drop-in boundary: context management, the batch launch, the host-buffer convenience path
// (H2D -> kernel -> D2H) and the Read-struct that mirrors brotli::Decompressor<R> (src/lib.rs:378-410,
// 2173-2193).  No decode work happens on the CPU anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/brotli_b200.h"
#include "bro_kernels.h"
#include "bro_status.h"

extern "C" const uint8_t bro_dictionary_blob[];   // csrc/dict_blob.c
#define BRO_DICT_BYTES 122784

struct bro_ctx {
    int device;
    int num_sms;
    int grid;                 // warp kernel, throughput build: persistent CTAs
    int grid_lat;             // warp kernel, latency build (fewer CTAs per SM, more registers)
    uint32_t num_warps;
    uint16_t* d_arena;        // warp kernel: worst-case arena per warp, for the first `arena_warps` warps of the grid
    uint32_t arena_warps;     // (grow-only, sized by the grids actually launched: a one-stream reader needs 8 warps = 10 MB,
                              // a full grid 4,736 warps = 5.9 GB)
    int grid_t;               // parse kernel: persistent CTAs
    uint32_t num_threads;
    uint16_t* d_arena_t;      // parse kernel: 64 KiB arena per thread
    uint16_t* d_roots; size_t roots_bytes;   // parse kernel: compact per-thread literal tables (bro_parse.h)
    int grid_c[2];            // copy kernel: persistent CTAs of its two shapes (bro_kernels.h)
    uint8_t* d_dict;
    uint32_t* d_counter;      // queue heads: [0] parse kernel, [1] warp kernel, [3] copy kernel; [2] retry count;
                              // [4..7] as two uint64: bytes moved / records executed by the copy kernel (last batch)
    uint32_t* d_order; size_t d_order_cap;   // size-class order of the batch | records per stream | completion queue (3 * cap)
    uint32_t* d_order_scratch;               // 512 counters
    BroRec* d_rec; size_t d_rec_cap;         // copy records (in records)
    cudaStream_t main;        // host-buffer path: copies in and kernels
    uint32_t* h_fault;        // host-buffer path: pinned, one fault word per chunk
    int host_chunks;          // host-buffer path: slices of a batch whose device -> host copy overlaps the decode of the next
    cudaStream_t side;        // host-buffer path: copies out (round 1: the copy kernel next to the parse kernel)
    cudaEvent_t ev_fork, ev_join;
    int overlap;              // 1: parse and copy kernels side by side (BRO_B200_OVERLAP=1); 0 (default): one after the other
                              // (always while timing is on, so that bro_ctx_last_kernel_ms reports each kernel alone)
    int timing;               // bro_ctx_set_timing: record CUDA events around the kernels of a batch
    cudaEvent_t ev[5];        // before ordering | before parse | before copy | before fused | after fused
    int ev_valid;             // the last batch recorded ev[] (two_phase: all five, else ev[3], ev[4])
    int ev_two_phase;
    uint64_t reserved_in;     // bro_ctx_reserve: the caller's bound on the compressed bytes of a batch (0 = none)
    int mode;                 // BRO_MODE_AUTO / WARP / TWOPHASE
    uint32_t twophase_threshold;   // AUTO: batches of at least this many streams take the two-phase path
    int quirks;
    long long watchdog;       // copy kernel: cycles to wait for one completion-queue slot (about half a minute)
    int parse_lanes;          // BRO_B200_PARSE_LANES=1..32 (tuning): streams per parse warp, 0 = by batch size
    int debug_no_parse;       // BRO_B200_DEBUG_NO_PARSE=1 (test-suite): the parse kernel is not launched, so the copy kernel's watchdog must fire
    uint64_t launches;
    char err[256];
    // grow-only device staging for the host-buffer path
    uint8_t* d_in; size_t d_in_cap;
    uint8_t* d_out; size_t d_out_cap;
    uint64_t* d_meta; size_t d_meta_cap;   // in_off | out_off | out_len | status
};

static int bro_fail(bro_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
    return BRO_ST_CudaError;
}
#define BRO_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bro_fail(ctx, e_, #call); } while (0)

extern "C" int bro_ctx_create(bro_ctx** out, int device) {
    if (!out) return BRO_ST_InvalidArgument;
    *out = NULL;
    bro_ctx* ctx = (bro_ctx*)calloc(1, sizeof(bro_ctx));
    if (!ctx) return BRO_ST_InvalidArgument;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) { free(ctx); return BRO_ST_CudaError; }
    if (cudaSetDevice(device) != cudaSuccess) { free(ctx); return BRO_ST_CudaError; }
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { free(ctx); return BRO_ST_CudaError; }
    ctx->num_sms = prop.multiProcessorCount;
    int per_sm2[2] = {0, 0}, per_sm = 0, per_sm_t = 0, per_sm_c[2] = {0, 0};
    if (bro_warp_kernel_occupancy(per_sm2) != 0 || (per_sm = per_sm2[0]) < 1 || per_sm2[1] < 1 ||
        bro_parse_kernel_occupancy(&per_sm_t) != 0 || per_sm_t < 1 ||
        bro_copy_kernel_occupancy(per_sm_c) != 0 || per_sm_c[0] < 1 || per_sm_c[1] < 1) { free(ctx); return BRO_ST_CudaError; }
    ctx->grid = ctx->num_sms * per_sm;
    ctx->grid_lat = ctx->num_sms * per_sm2[1];
    ctx->num_warps = (uint32_t)ctx->grid * (uint32_t)bro_warp_kernel_warps_per_cta();
    // The parse kernel is bound by the latency of its table look-ups, i.e. by how many streams' tables stay in L2:
    // BRO_B200_PARSE_BLOCKS caps its resident CTAs per SM (tuning knob).
    // Measured on B200 (profiles/r01_kernel_variants.md): 3 CTAs (384 streams) per SM is the optimum.
    int pb_want = 3;
    const char* pb = getenv("BRO_B200_PARSE_BLOCKS");
    if (pb && atoi(pb) >= 1) pb_want = atoi(pb);
    if (pb_want < per_sm_t) per_sm_t = pb_want;
    ctx->grid_t = ctx->num_sms * per_sm_t;
    ctx->num_threads = (uint32_t)ctx->grid_t * (uint32_t)bro_parse_kernel_block();
    ctx->grid_c[0] = ctx->num_sms * per_sm_c[0];
    ctx->grid_c[1] = ctx->num_sms * per_sm_c[1];
    ctx->mode = BRO_MODE_AUTO;
    // The two kernels of the two-phase path run one after the other: the parse kernel's CTA takes an SM's whole shared
    // memory (round 1 could run them side by side through the completion queue; it bought 5 % at best).
    ctx->overlap = 0;
    if (cudaStreamCreateWithFlags(&ctx->main, cudaStreamNonBlocking) != cudaSuccess) { ctx->main = NULL; cudaGetLastError(); }
    if (cudaHostAlloc((void**)&ctx->h_fault, 16 * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) { ctx->h_fault = NULL; cudaGetLastError(); }
    ctx->host_chunks = 4;
    { const char* hc = getenv("BRO_B200_HOST_CHUNKS"); if (hc && atoi(hc) >= 1 && atoi(hc) <= 16) ctx->host_chunks = atoi(hc); }
    { const char* pl = getenv("BRO_B200_PARSE_LANES"); if (pl && atoi(pl) >= 1 && atoi(pl) <= 32) ctx->parse_lanes = atoi(pl); }
    ctx->watchdog = 1ll << 36;
    { const char* dbg = getenv("BRO_B200_DEBUG_NO_PARSE"); if (dbg && dbg[0] == '1') { ctx->debug_no_parse = 1; ctx->watchdog = 1ll << 22; } }
    if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) { ctx->side = NULL; cudaGetLastError(); }
    // A warp per stream is the lowest latency per stream; the two-phase path (32 streams per warp in the entropy decode,
    // then copies at memory speed) has the higher throughput.  It pays once the fused kernel would need several waves
    // of its resident warps; measured on B200 with the headline streams (profiles/r01_kernel_variants.md) the break-even
    // is near 4 waves.
    ctx->twophase_threshold = 160u * (uint32_t)ctx->num_sms;      // 23,680 streams on a B200
    const char* env = getenv("BRO_B200_MODE");
    if (env && !strcmp(env, "warp")) ctx->mode = BRO_MODE_WARP;
    if (env && (!strcmp(env, "twophase") || !strcmp(env, "thread"))) ctx->mode = BRO_MODE_TWOPHASE;
    if (cudaMalloc(&ctx->d_dict, BRO_DICT_BYTES) != cudaSuccess ||
        cudaMalloc(&ctx->d_counter, 16 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&ctx->d_order_scratch, 512 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMemcpy(ctx->d_dict, bro_dictionary_blob, BRO_DICT_BYTES, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(ctx->d_arena); cudaFree(ctx->d_arena_t); cudaFree(ctx->d_dict); cudaFree(ctx->d_counter); cudaFree(ctx->d_order_scratch);
        free(ctx);
        return BRO_ST_CudaError;
    }
    *out = ctx;
    return BRO_ST_OK;
}

extern "C" void bro_ctx_destroy(bro_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_arena); cudaFree(ctx->d_arena_t); cudaFree(ctx->d_dict); cudaFree(ctx->d_counter);
    cudaFree(ctx->d_order); cudaFree(ctx->d_order_scratch); cudaFree(ctx->d_rec); cudaFree(ctx->d_roots);
    cudaFree(ctx->d_in); cudaFree(ctx->d_out); cudaFree(ctx->d_meta);
    for (int k = 0; k < 5; k++) if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->main) cudaStreamDestroy(ctx->main);
    if (ctx->h_fault) cudaFreeHost(ctx->h_fault);
    free(ctx);
}

extern "C" int bro_ctx_set_quirks(bro_ctx* ctx, int quirks) {
    if (!ctx || (quirks != 0 && quirks != 1)) return BRO_ST_InvalidArgument;
    ctx->quirks = quirks;
    return BRO_ST_OK;
}

extern "C" int bro_ctx_set_mode(bro_ctx* ctx, int mode) {
    if (!ctx || mode < BRO_MODE_AUTO || mode > BRO_MODE_TWOPHASE) return BRO_ST_InvalidArgument;
    ctx->mode = mode;
    return BRO_ST_OK;
}

extern "C" const char* bro_ctx_last_cuda_error(const bro_ctx* ctx) { return ctx ? ctx->err : "no context"; }
extern "C" uint64_t bro_ctx_launch_count(const bro_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint32_t bro_ctx_num_warps(const bro_ctx* ctx) { return ctx ? ctx->num_warps : 0; }

extern "C" int bro_ctx_set_timing(bro_ctx* ctx, int on) {
    if (!ctx) return BRO_ST_InvalidArgument;
    BRO_CUDA(ctx, cudaSetDevice(ctx->device));
    if (on && !ctx->ev[0]) for (int k = 0; k < 5; k++) BRO_CUDA(ctx, cudaEventCreate(&ctx->ev[k]));
    ctx->timing = on ? 1 : 0;
    ctx->ev_valid = 0;
    return BRO_ST_OK;
}

extern "C" int bro_ctx_last_kernel_ms(bro_ctx* ctx, float* ms4) {
    if (!ctx || !ms4) return BRO_ST_InvalidArgument;
    ms4[0] = ms4[1] = ms4[2] = ms4[3] = 0.0f;
    if (!ctx->ev_valid) return BRO_ST_InvalidArgument;
    BRO_CUDA(ctx, cudaEventSynchronize(ctx->ev[4]));
    if (ctx->ev_two_phase) for (int k = 0; k < 3; k++) BRO_CUDA(ctx, cudaEventElapsedTime(&ms4[k], ctx->ev[k], ctx->ev[k + 1]));
    BRO_CUDA(ctx, cudaEventElapsedTime(&ms4[3], ctx->ev[3], ctx->ev[4]));
    return BRO_ST_OK;
}

extern "C" int bro_ctx_last_batch_stats(bro_ctx* ctx, uint64_t* stats4) {
    if (!ctx || !stats4) return BRO_ST_InvalidArgument;
    uint32_t h[16];
    BRO_CUDA(ctx, cudaSetDevice(ctx->device));
    BRO_CUDA(ctx, cudaDeviceSynchronize());
    BRO_CUDA(ctx, cudaMemcpy(h, ctx->d_counter, sizeof(h), cudaMemcpyDeviceToHost));
    stats4[0] = (uint64_t)h[4] | ((uint64_t)h[5] << 32);     // bytes moved by copy records
    stats4[1] = (uint64_t)h[6] | ((uint64_t)h[7] << 32);     // copy records executed
    stats4[2] = h[2];                                        // streams handed to the fused kernel's retry pass
    stats4[3] = h[10];                                       // 1: AUTO's gate sent the whole batch to the fused kernel
    if (h[11]) {
        snprintf(ctx->err, sizeof(ctx->err), "copy kernel watchdog: a stream of the batch was never announced by the parse kernel");
        return BRO_ST_CudaError;
    }
    return BRO_ST_OK;
}

extern "C" int bro_ctx_reserve(bro_ctx* ctx, uint64_t total_in_bytes, uint32_t n_streams) {
    if (!ctx) return BRO_ST_InvalidArgument;
    (void)n_streams;
    ctx->reserved_in = total_in_bytes;
    return BRO_ST_OK;
}

// The device entry points run on the context's device whatever the caller's current device is, and leave the caller's
// choice as they found it.
struct BroDeviceGuard {
    int prev, ok;
    explicit BroDeviceGuard(int device) : prev(-1), ok(1) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = 0;
    }
    ~BroDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Table arenas of the warp kernels for a grid of `warps` warps (grow-only; nothing is in flight on a larger grid than
// the one that was allocated for, and growing synchronises the device)
static int bro_ensure_arena(bro_ctx* ctx, uint32_t warps) {
    if (warps <= ctx->arena_warps) return BRO_ST_OK;
    if (ctx->d_arena) { BRO_CUDA(ctx, cudaDeviceSynchronize()); BRO_CUDA(ctx, cudaFree(ctx->d_arena)); ctx->d_arena = NULL; ctx->arena_warps = 0; }
    uint32_t want = warps < 64u ? warps : (warps + 63u) & ~63u;
    if (want > ctx->num_warps) want = ctx->num_warps;
    BRO_CUDA(ctx, cudaMalloc(&ctx->d_arena, (size_t)want * bro_warp_kernel_arena_bytes()));
    ctx->arena_warps = want;
    return BRO_ST_OK;
}

static int bro_grow(bro_ctx* ctx, void** p, size_t* cap, size_t need, size_t elem) {
    if (*cap >= need) return BRO_ST_OK;
    if (*p) { BRO_CUDA(ctx, cudaFree(*p)); *p = NULL; *cap = 0; }
    size_t want = need + (need >> 3) + 1024;
    BRO_CUDA(ctx, cudaMalloc(p, want * elem));
    *cap = want;
    return BRO_ST_OK;
}

extern "C" int bro_batch_decode(bro_ctx* ctx, const uint8_t* d_in, const uint64_t* d_in_off, uint8_t* d_out,
                                const uint64_t* d_out_off, uint64_t* d_out_len, int32_t* d_status, uint32_t n,
                                void* stream) {
    if (!ctx) return BRO_ST_InvalidArgument;
    if (n == 0) return BRO_ST_OK;
    if (!d_in_off || !d_out_off || !d_out_len || !d_status) return BRO_ST_InvalidArgument;
    BroDeviceGuard guard(ctx->device);
    if (!guard.ok) return bro_fail(ctx, cudaErrorInvalidDevice, "cudaSetDevice(context's device)");
    cudaStream_t s = (cudaStream_t)stream;
    BRO_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, 16 * sizeof(uint32_t), s));
    BroLaunch p;
    memset(&p, 0, sizeof(p));
    p.in = d_in; p.in_off = d_in_off; p.out = d_out; p.out_off = d_out_off;
    p.out_len = d_out_len; p.status = d_status; p.n = n;
    p.dict = ctx->d_dict; p.quirk_spec = ctx->quirks;
    p.order = NULL; p.retry_count = ctx->d_counter + 2; p.retry_mode = 0;
    const uint32_t wpc = (uint32_t)bro_warp_kernel_warps_per_cta();
    // AUTO: the two-phase path for a large batch (160 streams per SM: whatever the streams are, the fused kernel would need
    // many waves), and from 32 streams per SM when the streams are not tiny (512 compressed bytes on average: below that a
    // stream is mostly headers and one warp per stream is as good) -- measured break-even on the headline streams: ~4,000
    uint64_t total_in = ctx->reserved_in;
    bool two_phase = ctx->mode == BRO_MODE_TWOPHASE || (ctx->mode == BRO_MODE_AUTO && n >= ctx->twophase_threshold);
    if (!two_phase && ctx->mode == BRO_MODE_AUTO && n >= 32u * (uint32_t)ctx->num_sms) {
        if (total_in == 0) {
            uint64_t ends[2];
            BRO_CUDA(ctx, cudaMemcpyAsync(&ends[0], d_in_off, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            BRO_CUDA(ctx, cudaMemcpyAsync(&ends[1], d_in_off + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            BRO_CUDA(ctx, cudaStreamSynchronize(s));
            if (ends[1] < ends[0]) return BRO_ST_InvalidArgument;
            total_in = ends[1] - ends[0] ? ends[1] - ends[0] : 1;
        }
        two_phase = total_in >= 512ull * n;
    }
    // The fused kernel has a throughput and a latency build (bro_kernels.cu).  On its own: by the number of streams.  Behind
    // the two-phase kernels both are launched and decide on the device (BroLaunch::fused_role).
    const uint32_t lat_warps = (uint32_t)ctx->grid_lat * wpc;
    const int latency = n <= lat_warps ? 1 : 0;
    int grid_w = ctx->grid, grid_l = ctx->grid_lat;
    if ((uint32_t)grid_w > (n + wpc - 1) / wpc) grid_w = (int)((n + wpc - 1) / wpc);
    if ((uint32_t)grid_l > (n + wpc - 1) / wpc) grid_l = (int)((n + wpc - 1) / wpc);
    { int st_a = bro_ensure_arena(ctx, (uint32_t)(grid_w > grid_l ? grid_w : grid_l) * wpc); if (st_a) return st_a; }       // before anything of this batch is in flight
    cudaError_t e;
    if (two_phase) {
        // PHASE ONE: one thread per stream (bro_parse_kernel), streams handed out by compressed-size class; literals and
        // dictionary words go into the slots, copies become records.  PHASE TWO: one warp per stream executes the
        // records (bro_copy_kernel).  Streams phase one cannot decode are left for the fused kernel's retry pass.
        int st;
        if (!ctx->d_arena_t) {   // 64 KiB per resident thread, allocated on first use
            BRO_CUDA(ctx, cudaMalloc(&ctx->d_arena_t, (size_t)ctx->num_threads * bro_parse_kernel_arena_bytes()));
            ctx->roots_bytes = (size_t)ctx->num_threads * bro_parse_kernel_roots_bytes();
            if (ctx->roots_bytes) {
                BRO_CUDA(ctx, cudaMalloc(&ctx->d_roots, ctx->roots_bytes));
            }
        }
        if ((st = bro_grow(ctx, (void**)&ctx->d_order, &ctx->d_order_cap, (size_t)n, 3 * sizeof(uint32_t)))) return st;
        // the record arena is sized from the compressed bytes of the batch: the caller's bound (bro_ctx_reserve), else
        // the two end offsets are read back (16 bytes, blocking on `s`)
        if (total_in == 0) {
            uint64_t ends[2];
            BRO_CUDA(ctx, cudaMemcpyAsync(&ends[0], d_in_off, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            BRO_CUDA(ctx, cudaMemcpyAsync(&ends[1], d_in_off + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            BRO_CUDA(ctx, cudaStreamSynchronize(s));
            if (ends[1] < ends[0]) return BRO_ST_InvalidArgument;
            total_in = ends[1] - ends[0];
        }
        const size_t rec_total = (size_t)(total_in >> 1) + 32u * (size_t)n;
        if ((st = bro_grow(ctx, (void**)&ctx->d_rec, &ctx->d_rec_cap, rec_total, sizeof(BroRec)))) return st;
        uint32_t* d_order = ctx->d_order;
        if (ctx->timing) BRO_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));
        p.nrec = ctx->d_order + ctx->d_order_cap;
        p.done_q = ctx->d_order + 2 * ctx->d_order_cap;
        p.done_tail = ctx->d_counter + 8;
        BRO_CUDA(ctx, cudaMemsetAsync(p.done_q, 0xff, (size_t)n * sizeof(uint32_t), s));
        p.rec = ctx->d_rec; p.rec_total = ctx->d_rec_cap;
        // AUTO: the ordering kernels also find the longest stream and decide on the device whether the batch is bound by
        // it; an explicit TWOPHASE is not second-guessed
        p.gate = ctx->mode == BRO_MODE_AUTO ? ctx->d_counter + 9 : NULL;
        e = (cudaError_t)bro_order_launch(d_in_off, n, d_order, ctx->d_order_scratch, p.gate, s);
        if (e != cudaSuccess) return bro_fail(ctx, e, "bro_order kernels launch");
        ctx->launches += 3;
        uint32_t tb = (uint32_t)bro_parse_kernel_block();
        // A batch of several waves (more streams than resident lanes): fewer warps per SM.  The parse kernel's throughput per SM
        // saturates at ten warps (a full wave of the headline streams takes 3.05 / 3.45 / 3.9 ms with 10 / 11 / 12 warps per SM,
        // i.e. 105 / 102 / 98 streams per ms), and what a batch costs is its waves, the last one nearly in full however few
        // streams it holds: cost(w) = wave time(w) x (waves + 0.35 x the empty part of the last wave), fitted to batches of
        // 60,000 ... 160,000 streams (profiles/r02_kernel_variants.md section 10).  The 100,000 headline streams are two even
        // waves of eleven warps: 6.76 ms against 7.07 ms with twelve and 7.54 ms with ten.
        {
            const uint32_t full = (uint32_t)ctx->grid_t * tb;                     // resident lanes with the full block
            if (n > full && !ctx->parse_lanes && tb == 384u) {
                static const float wave_ms[3] = {3.05f, 3.45f, 3.9f};
                float best = 0.f;
                uint32_t best_w = 12u;
                for (uint32_t w = 10u; w <= 12u; w++) {
                    const float waves = (float)n / (float)((uint32_t)ctx->grid_t * 32u * w);
                    const float frac = waves - (float)(uint32_t)waves;
                    const float cost = wave_ms[w - 10u] * (waves + (frac > 0.f ? 0.35f * (1.f - frac) : 0.f));
                    if (w == 10u || cost < best) { best = cost; best_w = w; }
                }
                tb = best_w * 32u;
            }
            const char* pw = getenv("BRO_B200_PARSE_WARPS");
            if (pw && atoi(pw) >= 2 && (uint32_t)atoi(pw) * 32u <= (uint32_t)bro_parse_kernel_block()) tb = (uint32_t)atoi(pw) * 32u;
        }
        // A batch smaller than the resident lanes is spread over all SMs: every warp takes fewer streams (a warp's lockstep
        // steps then serve fewer, busier lanes, and a stream's latency -- which is what a small batch costs -- drops)
        const uint32_t warps_total = (uint32_t)ctx->grid_t * (tb / 32u);
        uint32_t lanes = (n + warps_total - 1u) / warps_total;
        lanes = lanes < 4u ? 4u : lanes > 32u ? 32u : lanes;
        if (ctx->parse_lanes) lanes = (uint32_t)ctx->parse_lanes;
        p.lanes = lanes;
        int grid_t = ctx->grid_t;
        if ((uint32_t)grid_t > (n + lanes * (tb / 32u) - 1u) / (lanes * (tb / 32u))) grid_t = (int)((n + lanes * (tb / 32u) - 1u) / (lanes * (tb / 32u)));
        // the copy kernel's shape: the one whose warps take fewer streams one after the other, the throughput shape when both
        // take as many (any batch that fills the GPU several times over)
        const uint32_t cw0 = (uint32_t)bro_copy_kernel_warps_per_cta(0), cw1 = (uint32_t)bro_copy_kernel_warps_per_cta(1);
        const uint32_t warps0 = (uint32_t)ctx->grid_c[0] * cw0, warps1 = (uint32_t)ctx->grid_c[1] * cw1;
        int cshape = (n + warps1 - 1u) / warps1 < (n + warps0 - 1u) / warps0 && n < 8u * warps0 ? 1 : 0;
        { const char* cs = getenv("BRO_B200_COPY_SHAPE"); if (cs && (cs[0] == '0' || cs[0] == '1')) cshape = cs[0] - '0'; }
        const uint32_t cw = cshape ? cw1 : cw0;
        p.counter = ctx->d_counter + 3; p.order = NULL;
        p.copy_stats = (unsigned long long*)(ctx->d_counter + 4);
        const bool overlap = ctx->overlap && !ctx->timing && ctx->side;
        if (overlap) {
            // The copy kernel follows the parse kernel through the completion queue: launched on the side stream while
            // the parse kernel runs on the caller's, one CTA per SM (what fits next to the parse kernel's three).  The
            // parse kernel never waits for it, so whatever the block scheduler does there is no cycle to deadlock on.
            BRO_CUDA(ctx, cudaEventRecord(ctx->ev_fork, s));
            BRO_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        }
        p.arena = ctx->d_arena_t; p.counter = ctx->d_counter; p.order = d_order; p.roots = ctx->d_roots;
        if (ctx->timing) BRO_CUDA(ctx, cudaEventRecord(ctx->ev[1], s));
        e = ctx->debug_no_parse ? cudaSuccess : (cudaError_t)bro_parse_kernel_launch(&p, grid_t, (int)tb, s);
        if (e != cudaSuccess) return bro_fail(ctx, e, "bro_parse_kernel launch");
        ctx->launches += 1;
        p.counter = ctx->d_counter + 3; p.order = NULL;
        p.fault = ctx->d_counter + 11; p.watchdog = ctx->watchdog;
        int grid_c = ctx->grid_c[cshape];      // side by side, one CTA per SM fits next to the parse kernel; the others start as its CTAs exit
        if ((uint32_t)grid_c > (n + cw - 1) / cw) grid_c = (int)((n + cw - 1) / cw);
        if (ctx->timing) BRO_CUDA(ctx, cudaEventRecord(ctx->ev[2], s));
        e = (cudaError_t)bro_copy_kernel_launch(&p, grid_c, cshape, overlap ? ctx->side : s);
        if (e != cudaSuccess) return bro_fail(ctx, e, "bro_copy_kernel launch");
        ctx->launches += 1;
        if (overlap) {
            BRO_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side));
            BRO_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));
        }
        p.retry_mode = 1;
    }
    p.arena = ctx->d_arena; p.counter = ctx->d_counter + 1;
    p.order = two_phase ? ctx->d_order : NULL;       // size-class order of the batch (largest first) when it was computed
    if (ctx->timing) BRO_CUDA(ctx, cudaEventRecord(ctx->ev[3], s));
    if (two_phase) {
        p.fused_small = lat_warps;
        p.fused_role = 1;
        e = (cudaError_t)bro_warp_kernel_launch(&p, grid_l, 1, s);
        if (e == cudaSuccess && n > lat_warps) {       // (more streams than a latency wave could ever be retried)
            p.fused_role = 2;
            e = (cudaError_t)bro_warp_kernel_launch(&p, grid_w, 0, s);
            ctx->launches += 1;
        }
    } else e = (cudaError_t)bro_warp_kernel_launch(&p, latency ? grid_l : grid_w, latency, s);
    if (e != cudaSuccess) return bro_fail(ctx, e, "bro_decode_warp_kernel launch");
    ctx->launches += 1;
    if (ctx->timing) {
        BRO_CUDA(ctx, cudaEventRecord(ctx->ev[4], s));
        ctx->ev_valid = 1; ctx->ev_two_phase = two_phase ? 1 : 0;
    }
    return BRO_ST_OK;
}

static_assert(sizeof(bro_resume) == sizeof(BroResume), "bro_resume (include/brotli_b200.h) mirrors BroResume (bro_records.h)");

// The resumable decode: the fused warp-per-stream kernel from and to a resume point per stream (bro_kernels_resume.cu).
extern "C" int bro_batch_decode_resume(bro_ctx* ctx, const uint8_t* d_in, const uint64_t* d_in_off, uint8_t* d_out,
                                       const uint64_t* d_out_off, uint64_t* d_out_len, int32_t* d_status, bro_resume* d_resume,
                                       uint32_t n, void* stream) {
    if (!ctx) return BRO_ST_InvalidArgument;
    if (n == 0) return BRO_ST_OK;
    if (!d_in_off || !d_out_off || !d_out_len || !d_status || !d_resume) return BRO_ST_InvalidArgument;
    BroDeviceGuard guard(ctx->device);
    if (!guard.ok) return bro_fail(ctx, cudaErrorInvalidDevice, "cudaSetDevice(context's device)");
    cudaStream_t s = (cudaStream_t)stream;
    BRO_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, 16 * sizeof(uint32_t), s));
    BroLaunch p;
    memset(&p, 0, sizeof(p));
    p.in = d_in; p.in_off = d_in_off; p.out = d_out; p.out_off = d_out_off;
    p.out_len = d_out_len; p.status = d_status; p.n = n;
    p.dict = ctx->d_dict; p.quirk_spec = ctx->quirks;
    p.counter = ctx->d_counter + 1;
    p.resume = (BroResume*)d_resume;
    // the resume kernel has the warp kernel's CTA shape and arena layout: never more warps than the arena has shares
    const uint32_t wpc = (uint32_t)bro_resume_kernel_warps_per_cta();
    int grid = (int)(ctx->num_warps / wpc);
    if ((uint32_t)grid > (n + wpc - 1) / wpc) grid = (int)((n + wpc - 1) / wpc);
    { int st_a = bro_ensure_arena(ctx, (uint32_t)grid * wpc); if (st_a) return st_a; }
    p.arena = ctx->d_arena;
    cudaError_t e = (cudaError_t)bro_resume_kernel_launch(&p, grid, s);
    if (e != cudaSuccess) return bro_fail(ctx, e, "bro_decode_resume_kernel launch");
    ctx->launches += 1;
    return BRO_ST_OK;
}

extern "C" int bro_batch_sizes(bro_ctx* ctx, const uint8_t* d_in, const uint64_t* d_in_off, uint64_t* d_out_len,
                               int32_t* d_status, uint32_t n, void* stream) {
    if (!ctx) return BRO_ST_InvalidArgument;
    if (n == 0) return BRO_ST_OK;
    if (!d_in_off || !d_out_len || !d_status) return BRO_ST_InvalidArgument;
    BroDeviceGuard guard(ctx->device);
    if (!guard.ok) return bro_fail(ctx, cudaErrorInvalidDevice, "cudaSetDevice(context's device)");
    cudaStream_t s = (cudaStream_t)stream;
    int st;
    if (!ctx->d_arena_t) {
        BRO_CUDA(ctx, cudaMalloc(&ctx->d_arena_t, (size_t)ctx->num_threads * bro_parse_kernel_arena_bytes()));
        ctx->roots_bytes = (size_t)ctx->num_threads * bro_parse_kernel_roots_bytes();
        if (ctx->roots_bytes) BRO_CUDA(ctx, cudaMalloc(&ctx->d_roots, ctx->roots_bytes));
    }
    if ((st = bro_grow(ctx, (void**)&ctx->d_order, &ctx->d_order_cap, (size_t)n, 3 * sizeof(uint32_t)))) return st;
    BRO_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, 16 * sizeof(uint32_t), s));
    BroLaunch p;
    memset(&p, 0, sizeof(p));
    p.in = d_in; p.in_off = d_in_off; p.out_len = d_out_len; p.status = d_status; p.n = n;
    p.dict = ctx->d_dict; p.quirk_spec = ctx->quirks;
    p.retry_count = ctx->d_counter + 2;
    p.sizing = 1;
    cudaError_t e = (cudaError_t)bro_order_launch(d_in_off, n, ctx->d_order, ctx->d_order_scratch, NULL, s);
    if (e != cudaSuccess) return bro_fail(ctx, e, "bro_order kernels launch");
    ctx->launches += 3;
    const uint32_t tb = (uint32_t)bro_parse_kernel_block();
    int grid_t = ctx->grid_t;
    if ((uint32_t)grid_t > (n + tb - 1) / tb) grid_t = (int)((n + tb - 1) / tb);
    p.lanes = 32;
    p.arena = ctx->d_arena_t; p.counter = ctx->d_counter; p.order = ctx->d_order; p.roots = ctx->d_roots;
    e = (cudaError_t)bro_parse_kernel_launch(&p, grid_t, 0, s);
    if (e != cudaSuccess) return bro_fail(ctx, e, "bro_parse_kernel launch");
    ctx->launches += 1;
    e = (cudaError_t)bro_sizes_finish_launch(d_status, n, s);          // internal hand-over statuses -> BRO_SIZE_UNKNOWN
    if (e != cudaSuccess) return bro_fail(ctx, e, "bro_sizes_finish_kernel launch");
    ctx->launches += 1;
    return BRO_ST_OK;
}

extern "C" void bro_free(void* p) { free(p); }

static int bro_reserve(bro_ctx* ctx, void** p, size_t* cap, size_t need) {
    if (*cap >= need) return BRO_ST_OK;
    if (*p) { BRO_CUDA(ctx, cudaFree(*p)); *p = NULL; *cap = 0; }
    size_t want = need + (need >> 3) + 256;
    BRO_CUDA(ctx, cudaMalloc(p, want));
    *cap = want;
    return BRO_ST_OK;
}

// The end-to-end path of a host-language caller.  A batch's output is tens of times its input, so the device -> host copy
// is what the call costs (26 GB over PCIe for the headline batch); the decode is cut into up to host_chunks slices of
// streams and the copy of slice k runs on a second stream while slice k + 1 is decoded, so that the link never waits
// for a kernel.  Slices share the context's scratch: their kernels run one after the other on one stream.
extern "C" int bro_batch_decode_host(bro_ctx* ctx, const uint8_t* h_in, const uint64_t* h_in_off, uint8_t* h_out,
                                     const uint64_t* h_out_off, uint64_t* h_out_len, int32_t* h_status, uint32_t n) {
    if (!ctx) return BRO_ST_InvalidArgument;
    if (n == 0) return BRO_ST_OK;
    if (!h_in_off || !h_out_off || !h_out_len || !h_status) return BRO_ST_InvalidArgument;
    BroDeviceGuard guard(ctx->device);
    if (!guard.ok) return bro_fail(ctx, cudaErrorInvalidDevice, "cudaSetDevice(context's device)");
    const uint64_t in_lo = h_in_off[0], in_hi = h_in_off[n], out_lo = h_out_off[0], out_hi = h_out_off[n];
    if (in_hi < in_lo || out_hi < out_lo) return BRO_ST_InvalidArgument;
    const size_t in_bytes = (size_t)(in_hi - in_lo), out_bytes = (size_t)(out_hi - out_lo);
    const size_t off_bytes = (size_t)(n + 1) * sizeof(uint64_t);
    const size_t meta_bytes = 2 * off_bytes + (size_t)n * sizeof(uint64_t) + (size_t)n * sizeof(int32_t);
    int st;
    if ((st = bro_reserve(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, in_bytes + 16))) return st;
    if ((st = bro_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, out_bytes + 16))) return st;
    if ((st = bro_reserve(ctx, (void**)&ctx->d_meta, &ctx->d_meta_cap, meta_bytes))) return st;
    uint64_t* d_in_off = ctx->d_meta;
    uint64_t* d_out_off = d_in_off + (n + 1);
    uint64_t* d_out_len = d_out_off + (n + 1);
    int32_t* d_status = (int32_t*)(d_out_len + n);
    cudaStream_t s = ctx->main ? ctx->main : 0;
    cudaStream_t sc = (ctx->main && ctx->side && ctx->h_fault) ? ctx->side : s;      // copies out
    // slices: equal shares of the output bytes, at least 32 MB each (a small batch is one slice)
    uint32_t chunks = sc != s ? (uint32_t)ctx->host_chunks : 1u;
    while (chunks > 1u && (out_bytes / chunks < ((size_t)32 << 20) || n < 64u * chunks)) chunks--;
    uint32_t first[17];
    first[0] = 0;
    {
        uint32_t i = 0;
        for (uint32_t k = 1; k < chunks; k++) {
            const uint64_t want = out_lo + (uint64_t)((long double)out_bytes * k / chunks);
            while (i < n && h_out_off[i + 1] <= want) i++;
            first[k] = i > first[k - 1] ? i : first[k - 1];
        }
        first[chunks] = n;
    }
    BRO_CUDA(ctx, cudaMemcpyAsync(d_in_off, h_in_off, off_bytes, cudaMemcpyHostToDevice, s));
    BRO_CUDA(ctx, cudaMemcpyAsync(d_out_off, h_out_off, off_bytes, cudaMemcpyHostToDevice, s));
    const uint64_t reserved_saved = ctx->reserved_in;
    uint32_t fault_local = 0;
    for (uint32_t k = 0; k < chunks; k++) {
        const uint32_t a = first[k], b = first[k + 1];
        if (a == b) continue;
        const uint64_t ci_lo = h_in_off[a], ci_hi = h_in_off[b], co_lo = h_out_off[a], co_hi = h_out_off[b];
        if (ci_hi < ci_lo || co_hi < co_lo) { ctx->reserved_in = reserved_saved; return BRO_ST_InvalidArgument; }
        if (ci_hi > ci_lo) BRO_CUDA(ctx, cudaMemcpyAsync(ctx->d_in + (ci_lo - in_lo), h_in + ci_lo, (size_t)(ci_hi - ci_lo), cudaMemcpyHostToDevice, s));
        // offsets are relative to the caller's buffers; rebase the device pointers instead of rewriting the arrays
        ctx->reserved_in = ci_hi > ci_lo ? ci_hi - ci_lo : 1;          // known here: no read-back of the offsets
        st = bro_batch_decode(ctx, ctx->d_in - in_lo, d_in_off + a, ctx->d_out - out_lo, d_out_off + a, d_out_len + a, d_status + a, b - a, s);
        ctx->reserved_in = reserved_saved;
        if (st) return st;
        if (sc != s) {
            BRO_CUDA(ctx, cudaMemcpyAsync(ctx->h_fault + k, ctx->d_counter + 11, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            BRO_CUDA(ctx, cudaEventRecord(ctx->ev_fork, s));             // (one event: a later record only moves the wait point forward)
            BRO_CUDA(ctx, cudaStreamWaitEvent(sc, ctx->ev_fork, 0));
        } else {
            BRO_CUDA(ctx, cudaMemcpyAsync(&fault_local, ctx->d_counter + 11, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        }
        if (co_hi > co_lo) BRO_CUDA(ctx, cudaMemcpyAsync(h_out + co_lo, ctx->d_out + (co_lo - out_lo), (size_t)(co_hi - co_lo), cudaMemcpyDeviceToHost, sc));
    }
    BRO_CUDA(ctx, cudaMemcpyAsync(h_out_len, d_out_len, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    BRO_CUDA(ctx, cudaMemcpyAsync(h_status, d_status, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    BRO_CUDA(ctx, cudaStreamSynchronize(s));
    if (sc != s) BRO_CUDA(ctx, cudaStreamSynchronize(sc));
    uint32_t fault = fault_local;
    if (sc != s) for (uint32_t k = 0; k < chunks; k++) if (first[k] != first[k + 1]) fault |= ctx->h_fault[k];
    if (fault) {
        // the copy kernel gave up waiting for the parse kernel: statuses say OK for streams whose copies were never made
        snprintf(ctx->err, sizeof(ctx->err), "copy kernel watchdog: a stream of the batch was never announced by the parse kernel");
        for (uint32_t i = 0; i < n; i++) { h_status[i] = BRO_ST_CudaError; h_out_len[i] = 0; }
        return BRO_ST_CudaError;
    }
    return BRO_ST_OK;
}

extern "C" int bro_batch_decode_unsized_host(bro_ctx* ctx, const uint8_t* h_in, const uint64_t* h_in_off, uint32_t n,
                                             uint8_t** h_out, uint64_t* h_out_off, uint64_t* h_out_len, int32_t* h_status) {
    if (!ctx || !h_out) return BRO_ST_InvalidArgument;
    *h_out = NULL;
    if (n == 0) { if (h_out_off) h_out_off[0] = 0; return BRO_ST_OK; }
    if (!h_in_off || !h_out_off || !h_out_len || !h_status) return BRO_ST_InvalidArgument;
    BRO_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t in_lo = h_in_off[0], in_hi = h_in_off[n];
    if (in_hi < in_lo) return BRO_ST_InvalidArgument;
    const size_t in_bytes = (size_t)(in_hi - in_lo), off_bytes = (size_t)(n + 1) * sizeof(uint64_t);
    const size_t meta_bytes = 2 * off_bytes + (size_t)n * sizeof(uint64_t) + (size_t)n * sizeof(int32_t);
    int st;
    if ((st = bro_reserve(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, in_bytes + 16))) return st;
    if ((st = bro_reserve(ctx, (void**)&ctx->d_meta, &ctx->d_meta_cap, meta_bytes))) return st;
    uint64_t* d_in_off = ctx->d_meta;
    uint64_t* d_out_off = d_in_off + (n + 1);
    uint64_t* d_out_len = d_out_off + (n + 1);
    int32_t* d_status = (int32_t*)(d_out_len + n);
    cudaStream_t s = 0;
    try {
        // 1. measure (no output is written); the compressed batch stays on the device for the decode
        if (in_bytes) BRO_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, h_in + in_lo, in_bytes, cudaMemcpyHostToDevice, s));
        BRO_CUDA(ctx, cudaMemcpyAsync(d_in_off, h_in_off, off_bytes, cudaMemcpyHostToDevice, s));
        if ((st = bro_batch_sizes(ctx, ctx->d_in - in_lo, d_in_off, d_out_len, d_status, n, s))) return st;
        BRO_CUDA(ctx, cudaMemcpyAsync(h_out_len, d_out_len, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        BRO_CUDA(ctx, cudaMemcpyAsync(h_status, d_status, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        BRO_CUDA(ctx, cudaStreamSynchronize(s));
        // 2. slots: exact where the size is known, a guess (grown while the stream answers OutputTooSmall) elsewhere;
        //    invalid streams keep the status the measurement found and get an empty slot
        std::vector<uint64_t> cap(n), slot_off(n + 1);
        std::vector<uint32_t> todo;            // streams still to decode
        for (uint32_t i = 0; i < n; i++) {
            if (h_status[i] == BRO_ST_OK) { cap[i] = h_out_len[i]; todo.push_back(i); }
            else if (h_status[i] == BRO_ST_SizeUnknown) { cap[i] = 8 * (h_in_off[i + 1] - h_in_off[i]) + 4096; todo.push_back(i); }
            else { cap[i] = 0; h_out_len[i] = 0; }
        }
        std::vector<std::vector<uint8_t> > grown(n);     // final bytes of the streams that needed a second attempt
        std::vector<uint8_t> first;                      // slots of the first attempt
        std::vector<uint64_t> first_off(n + 1, 0);
        std::vector<int32_t> st_tmp;
        std::vector<uint64_t> len_tmp, off_in, off_out;
        for (int attempt = 0; !todo.empty(); attempt++) {
            const uint32_t m = (uint32_t)todo.size();
            off_in.assign(m + 1, 0); off_out.assign(m + 1, 0); len_tmp.assign(m, 0); st_tmp.assign(m, 0);
            // the sub-batch: compressed streams gathered on the host (first attempt: the batch as it is)
            std::vector<uint8_t> sub_in;
            const uint8_t* in_ptr = h_in;
            if (m == n) { for (uint32_t k = 0; k <= n; k++) off_in[k] = h_in_off[k]; }
            else {
                size_t tot = 0;
                for (uint32_t k = 0; k < m; k++) tot += (size_t)(h_in_off[todo[k] + 1] - h_in_off[todo[k]]);
                sub_in.resize(tot ? tot : 1);
                size_t o = 0;
                for (uint32_t k = 0; k < m; k++) {
                    const size_t len = (size_t)(h_in_off[todo[k] + 1] - h_in_off[todo[k]]);
                    if (len) memcpy(sub_in.data() + o, h_in + h_in_off[todo[k]], len);
                    off_in[k] = o; o += len;
                }
                off_in[m] = o;
                in_ptr = sub_in.data();
            }
            for (uint32_t k = 0; k < m; k++) off_out[k + 1] = off_out[k] + ((cap[todo[k]] + 15) & ~(uint64_t)15);
            std::vector<uint8_t> sub_out((size_t)off_out[m] ? (size_t)off_out[m] : 1);
            // bro_batch_decode_host takes slot i as [off[i], off[i+1]): the padding to 16 bytes belongs to the slot, which
            // is harmless (a stream that fits its exact size fits the padded slot)
            if ((st = bro_batch_decode_host(ctx, in_ptr, off_in.data(), sub_out.data(), off_out.data(), len_tmp.data(), st_tmp.data(), m))) return st;
            std::vector<uint32_t> again;
            for (uint32_t k = 0; k < m; k++) {
                const uint32_t i = todo[k];
                if (st_tmp[k] == BRO_ST_OutputTooSmall && cap[i] < 0xf0000000ull && attempt < 12) {
                    cap[i] = cap[i] * 4 < 0xf0000000ull ? cap[i] * 4 : 0xf0000000ull;
                    again.push_back(i);
                    continue;
                }
                h_status[i] = st_tmp[k];
                h_out_len[i] = st_tmp[k] == BRO_ST_OK ? len_tmp[k] : 0;
                if (attempt == 0) continue;            // first-attempt bytes stay in `first`
                grown[i].assign(sub_out.begin() + (size_t)off_out[k], sub_out.begin() + (size_t)off_out[k] + (size_t)h_out_len[i]);
            }
            if (attempt == 0) {
                first.swap(sub_out);
                for (uint32_t k = 0; k < m; k++) first_off[todo[k]] = off_out[k];
            }
            todo.swap(again);
        }
        // 3. one buffer, streams back to back (16-byte aligned starts)
        h_out_off[0] = 0;
        for (uint32_t i = 0; i < n; i++) h_out_off[i + 1] = h_out_off[i] + ((h_out_len[i] + 15) & ~(uint64_t)15);
        uint8_t* out = (uint8_t*)malloc((size_t)h_out_off[n] ? (size_t)h_out_off[n] : 1);
        if (!out) return BRO_ST_InvalidArgument;
        for (uint32_t i = 0; i < n; i++) {
            if (!h_out_len[i]) continue;
            const uint8_t* src = grown[i].empty() ? first.data() + (size_t)first_off[i] : grown[i].data();
            memcpy(out + h_out_off[i], src, (size_t)h_out_len[i]);
        }
        *h_out = out;
    } catch (...) {
        return BRO_ST_InvalidArgument;
    }
    return BRO_ST_OK;
}

// src/lib.rs:331-354 -- the strings are the observable error payload of the reference (io::Error description)
extern "C" const char* bro_status_description(int st) {
    switch (st) {
    case BRO_ST_OK: return "OK";
    case BRO_ST_CodeLengthsChecksum: return "Code length check sum did not add up in complex prefix code";
    case BRO_ST_ExpectedEndOfStream: return "Expected end-of-stream, but stream did not end";
    case BRO_ST_ExceededExpectedBytes: return "More uncompressed bytes than expected in meta-block";
    case BRO_ST_InvalidBlockCountCode: return "Encountered invalid value for block count code";
    case BRO_ST_InvalidBlockSwitchCommandCode: return "Encountered invalid value for block switch command code";
    case BRO_ST_InvalidLengthInStaticDictionary: return "Encountered invalid length in reference to static dictionary";
    case BRO_ST_InvalidMSkipLen: return "Most significant byte of MSKIPLEN was zero";
    case BRO_ST_InvalidSymbol: return "Encountered invalid symbol in prefix code";
    case BRO_ST_InvalidTransformId: return "Encountered invalid transform id in reference to static dictionary";
    case BRO_ST_InvalidNonPositiveDistance: return "Encountered invalid non-positive distance";
    case BRO_ST_LessThanTwoNonZeroCodeLengths: return "Encountered invalid complex prefix code with less than two non-zero codelengths";
    case BRO_ST_NoCodeLength: return "Encountered invalid complex prefix code with all zero codelengths";
    case BRO_ST_NonZeroFillBit: return "Enocuntered non-zero fill bit";
    case BRO_ST_NonZeroReservedBit: return "Enocuntered non-zero reserved bit";
    case BRO_ST_NonZeroTrailerBit: return "Enocuntered non-zero bit trailing the stream";
    case BRO_ST_NonZeroTrailerNibble: return "Enocuntered non-zero nibble trailing";
    case BRO_ST_ParseErrorContextMap: return "Error parsing context map";
    case BRO_ST_ParseErrorComplexPrefixCodeLengths: return "Error parsing code lengths for complex prefix code";
    case BRO_ST_ParseErrorDistanceCode: return "Error parsing DistanceCode";
    case BRO_ST_ParseErrorInsertAndCopyLength: return "Error parsing Insert And Copy Length";
    case BRO_ST_ParseErrorInsertLiterals: return "Error parsing Insert Literals";
    case BRO_ST_RingBufferError: return "Error accessing distance ring buffer";
    case BRO_ST_RunLengthExceededSizeOfContextMap: return "Run length excceeded declared length of context map";
    case BRO_ST_UnexpectedEOF: return "Encountered unexpected EOF";
    case BRO_ST_OutputTooSmall: return "output slot too small for the decoded stream";
    case BRO_ST_CudaError: return "CUDA error";
    case BRO_ST_PanicUppercaseZero: return "reference panics: uppercase_first on a dictionary word starting with 0x00";
    case BRO_ST_SizeUnknown: return "the size of the stream is known only after decoding it";
    case BRO_ST_InvalidArgument: return "invalid argument";
    default: return "unknown status";
    }
}

// ------------------------------------------------------------------------------------------------------
// The Read-struct: brotli::Decompressor<R: Read>
// ------------------------------------------------------------------------------------------------------
// Readers created without a context (what a drop-in `Decompressor::new(r)` does: the reference's test-suite creates
// dozens) share ONE lazily created context per device for the life of the process -- a few hundred KB of device memory
// until something is decoded, 10 MB of table arenas for one-stream batches -- and take turns on it.
#define BRO_MAX_DEVICES 64
static std::recursive_mutex g_default_mu;     // guards the table below AND every use of a default context (recursive: a
                                              // reader's callback may itself read from another default-context reader)
static bro_ctx* g_default_ctx[BRO_MAX_DEVICES];

static bro_ctx* bro_default_ctx(int* st) {    // call with g_default_mu held
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= BRO_MAX_DEVICES) { *st = BRO_ST_CudaError; return NULL; }
    if (!g_default_ctx[device]) {
        *st = bro_ctx_create(&g_default_ctx[device], device);
        if (*st) { g_default_ctx[device] = NULL; return NULL; }
    }
    *st = BRO_ST_OK;
    return g_default_ctx[device];
}

struct bro_reader {
    bro_ctx* ctx;
    bool shared_ctx;          // ctx is the process-wide default context of its device: used under g_default_mu
    bro_read_cb cb;
    void* user;
    bool decoded;             // whole-stream mode: the stream has been decoded; streaming mode: no call can add anything
    int status;
    std::vector<uint8_t> out; // decoded bytes not yet served (whole-stream mode: all of them)
    size_t served;
    // ---- streaming mode (bro_reader_new_streaming) ----
    bool streaming;
    size_t chunk;             // bytes asked of `cb` at a time
    std::vector<uint8_t> in;  // compressed bytes not yet consumed
    bool in_eof;
    size_t want;              // top `in` up to this many bytes before the next call
    BroResume ck;             // resume point (host copy)
    uint8_t* d_in; size_t d_in_cap;
    uint8_t* d_out[2]; size_t d_out_cap;      // ping-pong: the history moves to the other buffer between calls
    int cur;
    uint64_t* d_meta;         // in_off[2] | out_off[2] | out_len | status (+ pad) | BroResume
};

static bro_reader* bro_reader_alloc(bro_ctx* ctx, bro_read_cb cb, void* user) {
    if (!cb) return NULL;
    bro_reader* r = new (std::nothrow) bro_reader();
    if (!r) return NULL;
    r->ctx = ctx; r->shared_ctx = false; r->cb = cb; r->user = user;
    r->decoded = false; r->status = BRO_ST_OK; r->served = 0;
    r->streaming = false; r->chunk = 0; r->in_eof = false; r->want = 0;
    memset(&r->ck, 0, sizeof(r->ck));
    r->d_in = NULL; r->d_in_cap = 0; r->d_out[0] = r->d_out[1] = NULL; r->d_out_cap = 0; r->cur = 0; r->d_meta = NULL;
    return r;
}

extern "C" bro_reader* bro_reader_new(bro_ctx* ctx, bro_read_cb cb, void* user) {
    return bro_reader_alloc(ctx, cb, user);          // like Decompressor::new: no I/O yet (src/lib.rs:398-410)
}

extern "C" bro_reader* bro_reader_new_streaming(bro_ctx* ctx, bro_read_cb cb, void* user, size_t in_chunk) {
    bro_reader* r = bro_reader_alloc(ctx, cb, user);
    if (!r) return NULL;
    r->streaming = true;
    r->chunk = in_chunk ? in_chunk : (size_t)1 << 20;
    r->want = r->chunk;
    return r;
}

static void bro_reader_decode(bro_reader* r) {
    r->decoded = true;
    // drain R to end of input; an I/O error is indistinguishable from EOF for the reference's bit reader
    // (src/bitreader/mod.rs:78-82, 199), i.e. the stream is decoded as if it ended there
    std::vector<uint8_t> in;
    try {
        size_t chunk = 1 << 16;
        for (;;) {
            size_t old = in.size();
            in.resize(old + chunk);
            intptr_t got = r->cb(r->user, in.data() + old, chunk);
            if (got <= 0) { in.resize(old); break; }
            in.resize(old + (size_t)got);
            if (chunk < (1u << 24)) chunk <<= 1;
        }
        std::unique_lock<std::recursive_mutex> turn(g_default_mu, std::defer_lock);
        if (!r->ctx || r->shared_ctx) {
            turn.lock();
            int st = BRO_ST_OK;
            r->ctx = bro_default_ctx(&st);
            if (!r->ctx) { r->status = st; return; }
            r->shared_ctx = true;
        }
        size_t cap = in.size() * 6 + (1 << 16);
        for (;;) {
            r->out.resize(cap);
            uint64_t in_off[2] = {0, (uint64_t)in.size()}, out_off[2] = {0, (uint64_t)cap}, out_len = 0;
            int32_t status = 0;
            int st = bro_batch_decode_host(r->ctx, in.data(), in_off, r->out.data(), out_off, &out_len, &status, 1);
            if (st) { r->status = st; r->out.clear(); return; }
            if (status == BRO_ST_OutputTooSmall && cap < 0xf0000000ull) { cap = cap * 4 < 0xf0000000ull ? cap * 4 : 0xf0000000ull; continue; }
            r->status = status;
            // bytes produced before an error are not part of the contract: the reference loses the
            // output of the failing decompress() call, how much depends on the caller's read() sizes
            r->out.resize(status == BRO_ST_OK ? (size_t)out_len : 0);
            return;
        }
    } catch (...) {
        r->status = BRO_ST_InvalidArgument;
        r->out.clear();
    }
}

// ---- streaming mode: incremental input, bounded memory, resume at meta-block boundaries ----
//
// One step = top the input buffer up from R, one call of the resumable decode over (unconsumed input, history + room),
// then: the bytes in front of the new resume point are final and go to the caller; consumed input is dropped; the last
// min(window, bytes so far) bytes of output move to the front of the other output buffer as the next call's history.
// A call that ends inside a meta-block for lack of input (UnexpectedEOF while R has more) or room (OutputTooSmall) is
// repeated from the resume point with more of it -- twice as much when the call made no progress at all, which bounds
// both buffers by the largest meta-block (+ a window of history), not by the stream.

static int bro_reader_fail(bro_reader* r, int st) {
    r->status = st; r->decoded = true;
    return st;
}

#define BRO_RCUDA(r, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { bro_fail((r)->ctx, e_, #call); return bro_reader_fail(r, BRO_ST_CudaError); } } while (0)

static int bro_reader_out_reserve(bro_reader* r, size_t cap, size_t keep_from, size_t keep) {
    // both output buffers get capacity `cap`; bytes [keep_from, keep_from + keep) of the current one become [0, keep) of
    // the (new) current one
    uint8_t* nb[2] = {NULL, NULL};
    BRO_RCUDA(r, cudaMalloc(&nb[0], cap));
    if (cudaMalloc(&nb[1], cap) != cudaSuccess) { cudaFree(nb[0]); return bro_reader_fail(r, BRO_ST_CudaError); }
    if (keep) BRO_RCUDA(r, cudaMemcpy(nb[0], r->d_out[r->cur] + keep_from, keep, cudaMemcpyDeviceToDevice));
    cudaFree(r->d_out[0]); cudaFree(r->d_out[1]);
    r->d_out[0] = nb[0]; r->d_out[1] = nb[1]; r->d_out_cap = cap; r->cur = 0;
    return BRO_ST_OK;
}

// One step.  Returns BRO_ST_OK when the step may have added bytes to r->out or ended the stream (r->decoded).
static int bro_reader_step(bro_reader* r) {
    std::unique_lock<std::recursive_mutex> turn(g_default_mu, std::defer_lock);
    if (!r->ctx || r->shared_ctx) {
        turn.lock();
        int st = BRO_ST_OK;
        r->ctx = bro_default_ctx(&st);
        if (!r->ctx) return bro_reader_fail(r, st);
        r->shared_ctx = true;
    }
    cudaSetDevice(r->ctx->device);
    // top up the input
    while (!r->in_eof && r->in.size() < r->want) {
        const size_t old = r->in.size();
        r->in.resize(old + r->chunk);
        const intptr_t got = r->cb(r->user, r->in.data() + old, r->chunk);
        r->in.resize(old + (got > 0 ? (size_t)got : 0));
        if (got <= 0) r->in_eof = true;               // an I/O error ends the input (src/bitreader/mod.rs:78-82)
    }
    if (!r->d_meta) BRO_RCUDA(r, cudaMalloc(&r->d_meta, 8 * sizeof(uint64_t) + sizeof(BroResume)));
    if (!r->d_out[0]) {
        size_t cap = 4 * r->chunk;
        if (cap < ((size_t)1 << 16)) cap = (size_t)1 << 16;
        int st = bro_reader_out_reserve(r, cap, 0, 0);
        if (st) return st;
    }
    if (r->in.size() > r->d_in_cap) {
        cudaFree(r->d_in); r->d_in = NULL; r->d_in_cap = 0;
        const size_t cap = r->in.size() + r->in.size() / 2 + 64;
        BRO_RCUDA(r, cudaMalloc(&r->d_in, cap));
        r->d_in_cap = cap;
    }
    if (!r->d_in) { BRO_RCUDA(r, cudaMalloc(&r->d_in, 64)); r->d_in_cap = 64; }
    uint64_t meta[8] = {0, (uint64_t)r->in.size(), 0, (uint64_t)r->d_out_cap, 0, 0, 0, 0};
    BroResume* d_ck = (BroResume*)(r->d_meta + 8);
    if (!r->in.empty()) BRO_RCUDA(r, cudaMemcpy(r->d_in, r->in.data(), r->in.size(), cudaMemcpyHostToDevice));
    BRO_RCUDA(r, cudaMemcpy(r->d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice));
    BRO_RCUDA(r, cudaMemcpy(d_ck, &r->ck, sizeof(BroResume), cudaMemcpyHostToDevice));
    const BroResume before = r->ck;
    const uint32_t hist = r->ck.pos;
    int st = bro_batch_decode_resume(r->ctx, r->d_in, r->d_meta, r->d_out[r->cur], r->d_meta + 2, r->d_meta + 4, (int32_t*)(r->d_meta + 5),
                                     (bro_resume*)d_ck, 1, NULL);
    if (st) return bro_reader_fail(r, st);
    BRO_RCUDA(r, cudaMemcpy(meta, r->d_meta, sizeof(meta), cudaMemcpyDeviceToHost));
    BRO_RCUDA(r, cudaMemcpy(&r->ck, d_ck, sizeof(BroResume), cudaMemcpyDeviceToHost));
    const int32_t status = *(const int32_t*)&meta[5];
    const bool progress = r->ck.in_bits != before.in_bits || r->ck.pos != before.pos || r->ck.flags != before.flags;
    // bytes behind the last resume point belong to an unfinished meta-block and are decoded again
    const size_t final_pos = status == BRO_ST_OK ? (size_t)meta[4] : (size_t)r->ck.pos;
    if (final_pos > hist) {
        const size_t old = r->out.size();
        r->out.resize(old + (final_pos - hist));
        BRO_RCUDA(r, cudaMemcpy(r->out.data() + old, r->d_out[r->cur] + hist, final_pos - hist, cudaMemcpyDeviceToHost));
    }
    if (status == BRO_ST_OK) {
        // the device saw the end of ITS input behind the stream; bytes R still holds are trailing garbage (src/lib.rs:2160-2166)
        if (!r->in_eof) {
            uint8_t probe[64];
            const intptr_t got = r->cb(r->user, probe, sizeof(probe));
            if (got > 0) { r->status = BRO_ST_ExpectedEndOfStream; r->decoded = true; return BRO_ST_OK; }
            r->in_eof = true;
        }
        r->decoded = true;
        return BRO_ST_OK;
    }
    // consumed input goes; the history slides to the front of the other buffer
    const size_t drop = (size_t)(r->ck.in_bits >> 3);
    if (drop) r->in.erase(r->in.begin(), r->in.begin() + (drop < r->in.size() ? drop : r->in.size()));
    r->ck.in_bits &= 7u;
    size_t keep = 0;
    if (r->ck.flags & BRO_RESUME_HEADER) keep = r->ck.pos < r->ck.window ? r->ck.pos : r->ck.window;
    const size_t keep_from = (size_t)r->ck.pos - keep;
    if (status == BRO_ST_OutputTooSmall && !progress) {
        if (r->d_out_cap >= ((size_t)1 << 31)) { r->status = status; r->decoded = true; return BRO_ST_OK; }
        int st2 = bro_reader_out_reserve(r, 2 * r->d_out_cap, keep_from, keep);
        if (st2) return st2;
    } else if (keep_from != 0) {
        if (keep) BRO_RCUDA(r, cudaMemcpy(r->d_out[1 - r->cur], r->d_out[r->cur] + keep_from, keep, cudaMemcpyDeviceToDevice));
        r->cur = 1 - r->cur;
    }
    r->ck.pos = (uint32_t)keep;
    if (status == BRO_ST_UnexpectedEOF && !r->in_eof) {
        r->want = progress ? r->in.size() + 1 : 2 * r->in.size() + 1;     // no progress: this meta-block needs more at once
        if (r->want < r->chunk) r->want = r->chunk;
        return BRO_ST_OK;
    }
    if (status == BRO_ST_OutputTooSmall) return BRO_ST_OK;
    r->status = status;                               // an invalid stream (or one that ends early): final
    r->decoded = true;
    return BRO_ST_OK;
}

extern "C" intptr_t bro_reader_read(bro_reader* r, uint8_t* buf, size_t len) {
    if (!r || (!buf && len)) return -(intptr_t)BRO_ST_InvalidArgument;
    if (r->streaming) {
        // like Read::read of the reference (src/lib.rs:2174-2192): fill `buf` while the stream has data; bytes decoded
        // before an error are delivered first, the error by the read that finds nothing in front of it
        size_t done = 0;
        try {
            for (;;) {
                const size_t left = r->out.size() - r->served;
                const size_t k = len - done < left ? len - done : left;
                if (k) { memcpy(buf + done, r->out.data() + r->served, k); r->served += k; done += k; }
                if (r->served == r->out.size()) { r->out.clear(); r->served = 0; }
                if (done == len || r->decoded) break;
                if (bro_reader_step(r) != BRO_ST_OK) break;
            }
        } catch (...) {
            bro_reader_fail(r, BRO_ST_InvalidArgument);
        }
        if (done) return (intptr_t)done;
        if (r->status != BRO_ST_OK) return -(intptr_t)r->status;
        return 0;
    }
    if (!r->decoded) bro_reader_decode(r);
    if (r->status != BRO_ST_OK) return -(intptr_t)r->status;   // io::Error(InvalidData, description), src/lib.rs:2177
    size_t left = r->out.size() - r->served;
    size_t k = len < left ? len : left;
    if (k) memcpy(buf, r->out.data() + r->served, k);
    r->served += k;
    return (intptr_t)k;                                        // 0 = end of stream, repeatable
}

extern "C" int bro_reader_status(const bro_reader* r) { return r ? r->status : BRO_ST_InvalidArgument; }

extern "C" void bro_reader_free(bro_reader* r) {
    if (!r) return;
    if (r->d_in || r->d_out[0] || r->d_meta) {
        if (r->ctx) cudaSetDevice(r->ctx->device);
        cudaFree(r->d_in); cudaFree(r->d_out[0]); cudaFree(r->d_out[1]); cudaFree(r->d_meta);
    }
    delete r;                                          // (a default context lives as long as the process)
}

// ------------------------------------------------------------------------------------------------------
// Multi-GPU in one process: one context per device, the batch split by stream into one
// contiguous range per device with (nearly) equal work, every range decoded by its own host thread through
// bro_batch_decode_host.  Streams are independent: no byte crosses between GPUs.
// ------------------------------------------------------------------------------------------------------
struct bro_mg {
    int n;
    bro_ctx** ctx;
};

extern "C" int bro_mg_create(bro_mg** out, int ngpus) {
    if (!out) return BRO_ST_InvalidArgument;
    *out = NULL;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) return BRO_ST_CudaError;
    if (ngpus <= 0) ngpus = have;
    if (ngpus > have) return BRO_ST_InvalidArgument;
    int prev = -1;
    cudaGetDevice(&prev);
    bro_mg* mg = (bro_mg*)calloc(1, sizeof(bro_mg));
    if (!mg) return BRO_ST_InvalidArgument;
    mg->ctx = (bro_ctx**)calloc((size_t)ngpus, sizeof(bro_ctx*));
    if (!mg->ctx) { free(mg); return BRO_ST_InvalidArgument; }
    mg->n = ngpus;
    int st = BRO_ST_OK;
    for (int k = 0; k < ngpus && !st; k++) st = bro_ctx_create(&mg->ctx[k], k);
    if (prev >= 0) cudaSetDevice(prev);
    if (st) { bro_mg_destroy(mg); return st; }
    *out = mg;
    return BRO_ST_OK;
}

extern "C" void bro_mg_destroy(bro_mg* mg) {
    if (!mg) return;
    int prev = -1;
    cudaGetDevice(&prev);
    for (int k = 0; k < mg->n; k++) bro_ctx_destroy(mg->ctx[k]);
    if (prev >= 0) cudaSetDevice(prev);
    free(mg->ctx);
    free(mg);
}

extern "C" int bro_mg_device_count(const bro_mg* mg) { return mg ? mg->n : 0; }
extern "C" bro_ctx* bro_mg_ctx(bro_mg* mg, int k) { return (mg && k >= 0 && k < mg->n) ? mg->ctx[k] : NULL; }

// first[k] .. first[k + 1]: the streams of device k.  Work of a stream = compressed bytes + slot bytes (what crosses
// PCIe for it and, to first order, what its decode costs); the cut points are where the running sum passes k / ngpus of
// the total.
extern "C" int bro_mg_partition(const uint64_t* h_in_off, const uint64_t* h_out_off, uint32_t n, int ngpus, uint32_t* first) {
    if (!h_in_off || !h_out_off || !first || ngpus < 1) return BRO_ST_InvalidArgument;
    const long double total = (long double)(h_in_off[n] - h_in_off[0]) + (long double)(h_out_off[n] - h_out_off[0]);
    uint32_t i = 0;
    first[0] = 0;
    for (int k = 1; k < ngpus; k++) {
        const long double want = total * k / ngpus;
        while (i < n && (long double)(h_in_off[i + 1] - h_in_off[0]) + (long double)(h_out_off[i + 1] - h_out_off[0]) <= want) i++;
        first[k] = i;
    }
    first[ngpus] = n;
    return BRO_ST_OK;
}

extern "C" int bro_mg_decode_host(bro_mg* mg, const uint8_t* h_in, const uint64_t* h_in_off, uint8_t* h_out,
                                  const uint64_t* h_out_off, uint64_t* h_out_len, int32_t* h_status, uint32_t n) {
    if (!mg) return BRO_ST_InvalidArgument;
    if (n == 0) return BRO_ST_OK;
    if (!h_in_off || !h_out_off || !h_out_len || !h_status) return BRO_ST_InvalidArgument;
    try {
        std::vector<uint32_t> first((size_t)mg->n + 1);
        int st = bro_mg_partition(h_in_off, h_out_off, n, mg->n, first.data());
        if (st) return st;
        std::vector<int> rc((size_t)mg->n, BRO_ST_OK);
        std::vector<std::thread> workers;
        for (int k = 0; k < mg->n; k++) {
            const uint32_t a = first[k], b = first[k + 1];
            if (a == b) continue;
            // bro_batch_decode_host takes offsets relative to the caller's buffers: a range of the batch is the same
            // buffers with the offset arrays advanced
            workers.emplace_back([=, &rc]() {
                rc[k] = bro_batch_decode_host(mg->ctx[k], h_in, h_in_off + a, h_out, h_out_off + a, h_out_len + a, h_status + a, b - a);
            });
        }
        for (auto& w : workers) w.join();
        for (int k = 0; k < mg->n; k++) if (rc[k]) return rc[k];
    } catch (...) {
        return BRO_ST_InvalidArgument;
    }
    return BRO_ST_OK;
}
