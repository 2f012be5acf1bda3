// bro_kernels.cu -- the WARP-PER-STREAM decode kernel for sm_100a and its launcher.
//
// One persistent warp per stream at a time: warps pull stream indices from a global counter until the batch
// is exhausted, so streams of very different sizes (1 B .. hundreds of KB, 0 .. 65,537 meta-blocks) balance
// themselves.  The decoder itself is bro_decoder_core.h (32-lane mode); this file provides the per-warp resources
// (shared-memory scratch, a worst-case table arena in HBM) and the grid.  It serves small batches (lowest latency
// per stream) and, as the retry pass of the two-phase path, the streams the parse kernel hands over (literal context
// modelling, tables larger than a thread arena, more copy records than the stream's share).
#include <cuda_runtime.h>
#include <stdint.h>

#include "bro_decoder_core.h"
#include "bro_kernels.h"

#ifndef BRO_MIN_BLOCKS
#define BRO_MIN_BLOCKS 4
#endif

// A full warp per stream has worst-case arenas (it is also the retry kernel); experimental sub-warp groups
// (-DBRO_GROUP_W=16|8) get 128 KiB each.
#define BRO_GROUP_ARENA_U16 (BRO_W == 32u ? BRO_ARENA_U16_MAX : 65536u)
#define BRO_GROUPS_PER_WARP (32u / BRO_W)

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, BRO_MIN_BLOCKS) bro_decode_warp_kernel(BroLaunch p) {
    __shared__ BroScratch scratch[WARPS * BRO_GROUPS_PER_WARP];
    const unsigned warp = threadIdx.x / BRO_W, lane = bro_lane();     // "warp" = group of BRO_W lanes
    const unsigned gwarp = blockIdx.x * (WARPS * BRO_GROUPS_PER_WARP) + warp;
    // retry pass of the two-phase path: only the streams the parse kernel handed over -- unless AUTO's gate sent the
    // whole batch here
    const bool retry = p.retry_mode && !(p.gate && p.gate[1]);
    if (retry && *p.retry_count == 0u) return;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) {
            i = atomicAdd(p.counter, 1u);
            // after the ordering kernels (two-phase path, AUTO's gate): longest streams first, so that the batch does not
            // end on a 900 KB stream that was handed out last
            if (p.order && i < p.n) i = p.order[i];
        }
        i = bro_shfl(i, 0);
        if (i >= p.n) break;
        if (retry && !BRO_ST_IS_RETRY(p.status[i])) continue;
        const uint64_t in_b = p.in_off[i], in_e = p.in_off[i + 1];
        const uint64_t out_b = p.out_off[i], out_e = p.out_off[i + 1];
        BroDec d;
        d.sc = &scratch[warp];
        d.arena = p.arena + (size_t)gwarp * BRO_GROUP_ARENA_U16;
        d.arena_cap = BRO_GROUP_ARENA_U16;
        d.arena_base = 0;
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
        bro_syncwarp();
        if (lane == 0) {
            p.status[i] = st;
            p.out_len[i] = d.pos;
        }
    }
}

#define BRO_WARPS_PER_CTA 8

extern "C" int bro_warp_kernel_occupancy(int* blocks_per_sm) {
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, bro_decode_warp_kernel<BRO_WARPS_PER_CTA>,
                                                              BRO_WARPS_PER_CTA * 32, 0);
}

extern "C" int bro_warp_kernel_warps_per_cta() { return BRO_WARPS_PER_CTA * BRO_GROUPS_PER_WARP; }
extern "C" size_t bro_warp_kernel_arena_bytes() { return 2u * (size_t)BRO_GROUP_ARENA_U16; }

extern "C" int bro_warp_kernel_launch(const BroLaunch* p, int grid, cudaStream_t stream) {
    (void)cudaGetLastError();      // a stale error of another library in this process (it is per thread) is not this launch's
    bro_decode_warp_kernel<BRO_WARPS_PER_CTA><<<grid, BRO_WARPS_PER_CTA * 32, 0, stream>>>(*p);
    return (int)cudaGetLastError();
}
