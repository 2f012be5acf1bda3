// bro_kernels.cu -- the batched Brotli decode kernel for sm_100a and its launcher.
//
// One persistent warp per stream at a time: warps pull stream indices from a global counter until the batch
// is exhausted, so streams of very different sizes (1 B .. hundreds of KB, 0 .. 65,537 meta-blocks) balance
// themselves.  The decoder itself is bro_decoder_core.h; this file only provides the per-warp resources
// (shared-memory scratch, the table arena in HBM) and the grid.
#include <cuda_runtime.h>
#include <stdint.h>

#include "bro_decoder_core.h"
#include "bro_kernels.h"

#ifndef BRO_MIN_BLOCKS
#define BRO_MIN_BLOCKS 4
#endif

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, BRO_MIN_BLOCKS) bro_decode_kernel(BroLaunch p) {
    __shared__ BroScratch scratch[WARPS];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned gwarp = blockIdx.x * WARPS + warp;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(p.counter, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= p.n) break;
        const uint64_t in_b = p.in_off[i], in_e = p.in_off[i + 1];
        const uint64_t out_b = p.out_off[i], out_e = p.out_off[i + 1];
        BroDec d;
        d.sc = &scratch[warp];
        d.arena = p.arena + (size_t)gwarp * BRO_ARENA_U16;
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
        __syncwarp();
        if (lane == 0) {
            p.status[i] = st;
            p.out_len[i] = d.pos;
        }
    }
}

#define BRO_WARPS_PER_CTA 8

extern "C" int bro_kernel_occupancy(int* blocks_per_sm) {
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, bro_decode_kernel<BRO_WARPS_PER_CTA>,
                                                              BRO_WARPS_PER_CTA * 32, 0);
}

extern "C" int bro_kernel_warps_per_cta() { return BRO_WARPS_PER_CTA; }
extern "C" size_t bro_kernel_arena_bytes_per_warp() { return BRO_ARENA_BYTES; }

extern "C" int bro_kernel_launch(const BroLaunch* p, int grid, cudaStream_t stream) {
    bro_decode_kernel<BRO_WARPS_PER_CTA><<<grid, BRO_WARPS_PER_CTA * 32, 0, stream>>>(*p);
    return (int)cudaGetLastError();
}
