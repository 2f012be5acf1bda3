// bro_kernels_thread.cu -- the THREAD-PER-STREAM decode kernel for sm_100a, and the size-class ordering kernels.
//
// Brotli's entropy decode is serial inside a stream, so a warp that decodes ONE stream spends 31/32 of its issue
// slots on redundant lanes (ncu: bro_decode_warp_kernel is issue-bound at ~390 k warp instructions per 266 KB
// stream).  Here every THREAD owns a stream and runs the same decoder (bro_decoder_core.h with a 1-lane "warp"):
// one warp instruction advances 32 streams.  Divergence is what it costs, so streams are handed out in size-class
// order (similar streams share a warp; in replicated batches, replicas of the same stream do).  Each thread has a
// 64 KiB table arena in HBM (L1/L2 resident while hot); a meta-block that needs more is left with status
// ArenaTooSmall for the warp kernel's retry pass.
#include <cuda_runtime.h>
#include <stdint.h>

#define BRO_THREAD_MODE 1
#include "bro_decoder_core.h"
#include "bro_kernels.h"

#ifndef BRO_THREAD_BLOCK
#define BRO_THREAD_BLOCK 128
#endif
#ifndef BRO_THREAD_MIN_BLOCKS
#define BRO_THREAD_MIN_BLOCKS 4
#endif
#define BRO_SCRATCH_U16 ((sizeof(BroScratch) / 2u + 7u) & ~7u)
#define BRO_THREAD_ARENA_STRIDE_U16 (BRO_THREAD_ARENA_U16 + 704u)

__global__ void __launch_bounds__(BRO_THREAD_BLOCK, BRO_THREAD_MIN_BLOCKS) bro_decode_thread_kernel(BroLaunch p) {
    const unsigned t = blockIdx.x * BRO_THREAD_BLOCK + threadIdx.x;
    // the stride is an odd number of 128-byte lines, so that the threads' root tables do not pile into a few L1 sets
    uint16_t* arena = p.arena + (size_t)t * BRO_THREAD_ARENA_STRIDE_U16;
    for (;;) {
        uint32_t k = atomicAdd(p.counter, 1u);
        if (k >= p.n) break;
        const uint32_t i = p.order ? p.order[k] : k;
        const uint64_t in_b = p.in_off[i], in_e = p.in_off[i + 1];
        const uint64_t out_b = p.out_off[i], out_e = p.out_off[i + 1];
        BroDec d;
        d.sc = (BroScratch*)arena;
        d.arena = arena;
        d.arena_cap = BRO_THREAD_ARENA_U16;
        d.arena_base = BRO_SCRATCH_U16;
        d.dict = p.dict;
        d.out = p.out + out_b;
        uint64_t cap = out_e - out_b;
        d.cap = cap > BRO_MAX_SLOT ? (uint32_t)BRO_MAX_SLOT : (uint32_t)cap;
        d.pos = 0;
        d.p1 = 0; d.p2 = 0;
        d.d0 = 4; d.d1 = 11; d.d2 = 15; d.d3 = 16;   // src/lib.rs:407-408
        d.quirk_spec = p.quirk_spec;
        bro_bits_init(d.in, p.in + in_b, p.in + in_e);
        int st = bro_decode_stream(d);
        if (st == BRO_ST_ArenaTooSmall) atomicAdd(p.retry_count, 1u);
        p.status[i] = st;
        p.out_len[i] = d.pos;
    }
}

// ---- size-class ordering: 256 classes (8 per power of two of the compressed size), largest class first ----
__device__ __forceinline__ uint32_t bro_size_class(uint64_t len) {
    uint32_t l = len > 0xffffffffull ? 0xffffffffu : (uint32_t)len;
    uint32_t b = l ? 31u - (uint32_t)__clz(l) : 0u;
    uint32_t sub = b >= 3u ? (l >> (b - 3u)) & 7u : 0u;
    return 255u - (b * 8u + sub);
}

__global__ void bro_order_hist_kernel(const uint64_t* in_off, uint32_t n, uint32_t* hist) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&hist[bro_size_class(in_off[i + 1] - in_off[i])], 1u);
}

__global__ void bro_order_scan_kernel(const uint32_t* hist, uint32_t* cursor) {
    __shared__ uint32_t s[256];
    uint32_t t = threadIdx.x;
    s[t] = hist[t];
    __syncthreads();
    if (t == 0) {
        uint32_t run = 0;
        for (int k = 0; k < 256; k++) { uint32_t c = s[k]; s[k] = run; run += c; }
    }
    __syncthreads();
    cursor[t] = s[t];
}

__global__ void bro_order_scatter_kernel(const uint64_t* in_off, uint32_t n, uint32_t* cursor, uint32_t* order) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[atomicAdd(&cursor[bro_size_class(in_off[i + 1] - in_off[i])], 1u)] = i;
}

extern "C" int bro_order_launch(const uint64_t* in_off, uint32_t n, uint32_t* order, uint32_t* scratch, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, 512 * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return (int)e;
    uint32_t blocks = (n + 255u) / 256u;
    bro_order_hist_kernel<<<blocks, 256, 0, stream>>>(in_off, n, scratch);
    bro_order_scan_kernel<<<1, 256, 0, stream>>>(scratch, scratch + 256);
    bro_order_scatter_kernel<<<blocks, 256, 0, stream>>>(in_off, n, scratch + 256, order);
    return (int)cudaGetLastError();
}

extern "C" int bro_thread_kernel_occupancy(int* blocks_per_sm) {
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, bro_decode_thread_kernel, BRO_THREAD_BLOCK, 0);
}
extern "C" int bro_thread_kernel_block() { return BRO_THREAD_BLOCK; }
extern "C" size_t bro_thread_kernel_arena_bytes() { return 2u * (size_t)BRO_THREAD_ARENA_STRIDE_U16; }

extern "C" int bro_thread_kernel_launch(const BroLaunch* p, int grid, cudaStream_t stream) {
    bro_decode_thread_kernel<<<grid, BRO_THREAD_BLOCK, 0, stream>>>(*p);
    return (int)cudaGetLastError();
}
