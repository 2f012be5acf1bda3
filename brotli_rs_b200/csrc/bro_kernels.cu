// bro_kernels.cu -- the WARP-PER-STREAM decode kernel for sm_100a and its launcher.
//
// One persistent warp per stream at a time: warps pull stream indices from a global counter until the batch
// is exhausted, so streams of very different sizes (1 B .. hundreds of KB, 0 .. 65,537 meta-blocks) balance
// themselves.  The decoder itself is bro_decoder_core.h (32-lane mode); this file provides the per-warp resources
// (shared-memory scratch, a worst-case table arena in HBM) and the grid.  It serves small batches (lowest latency
// per stream) and, as the retry pass of the two-phase path, the streams the parse kernel hands over (literal context
// modelling, tables larger than a thread arena, more copy records than the stream's share).
#if !defined(BRO_WARPSIM)   /* (BRO_WARPSIM: this kernel compiled for the host, 32 lanes as fibers -- CPU test-suite only, bro_warpsim.cpp) */
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "bro_decoder_core.h"
#include "bro_kernels.h"

// Two builds of the same kernel (measured on B200, profiles/r02_kernel_variants.md):
//   throughput: 3 CTAs of 8 warps per SM, 80 registers -- batches of a few thousand to a few ten thousand streams
//               (10,000 x quickfox: 0.73 ms; 4 CTAs / 64 registers: 0.81 ms, 2 CTAs: 0.89 ms);
//   latency:    2 CTAs per SM, 128 registers, no spills -- whenever a single stream's decode time is what the call costs:
//               one wave of streams or less, the retry pass of the two-phase path, a batch AUTO's gate found bound by its
//               longest stream (alice29 alone: 24 ms against 33 ms; corpus x 1000: 300 ms against 343 ms)
#ifndef BRO_MIN_BLOCKS
#define BRO_MIN_BLOCKS 3
#endif
#ifndef BRO_MIN_BLOCKS_LATENCY
#define BRO_MIN_BLOCKS_LATENCY 2
#endif
// BRO_DICT_SMEM=1 (measured variant, not the product: profiles/r02_kernel_variants.md): the 122,784-byte static dictionary
// image in shared memory, one CTA of BRO_WARPS_PER_CTA (14) warps per SM -- what BASELINE's north_star suggests.  The
// product keeps the image in HBM / L2 (it is L1-resident wherever it is used) and the shared memory for 24 warps per SM.
#ifndef BRO_DICT_SMEM
#define BRO_DICT_SMEM 0
#endif
#define BRO_DICT_IMAGE_BYTES 122784u

// A full warp per stream has worst-case arenas (it is also the retry kernel); experimental sub-warp groups
// (-DBRO_GROUP_W=16|8) get 128 KiB each.
#define BRO_GROUP_ARENA_U16 (BRO_W == 32u ? BRO_ARENA_U16_MAX : 65536u)
#define BRO_GROUPS_PER_WARP (32u / BRO_W)

#ifndef BRO_WARPS_PER_CTA
#define BRO_WARPS_PER_CTA 8
#endif
#define BRO_WARP_KERNEL_SMEM_BASE (((BRO_WARPS_PER_CTA * BRO_GROUPS_PER_WARP * sizeof(BroScratch) + 15u) & ~(size_t)15) + (BRO_DICT_SMEM ? BRO_DICT_IMAGE_BYTES : 0u))
#define BRO_WARP_KERNEL_SMEM_LATENCY (BRO_WARP_KERNEL_SMEM_BASE + 2048u * BRO_WARPS_PER_CTA)
// (a variant build whose two builds are one and the same instance gets the larger launch for both)
#define BRO_WARP_KERNEL_SMEM (BRO_MIN_BLOCKS == BRO_MIN_BLOCKS_LATENCY ? BRO_WARP_KERNEL_SMEM_LATENCY : BRO_WARP_KERNEL_SMEM_BASE)

template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) bro_decode_warp_kernel(BroLaunch p) {
    // one BroScratch per warp (6.7 KB with the general loop's on-chip tables: dynamic shared memory, 4 CTAs x 8 warps = 214 KB per SM)
#if defined(BRO_WARPSIM)
    uint8_t* const bro_smem_raw = ws_dynamic_smem;
#else
    extern __shared__ __align__(16) uint8_t bro_smem_raw[];
#endif
    BroScratch* const scratch = (BroScratch*)bro_smem_raw;
#if BRO_DICT_SMEM
    uint8_t* const s_dict = bro_smem_raw + ((WARPS * BRO_GROUPS_PER_WARP * sizeof(BroScratch) + 15u) & ~(size_t)15);
    for (unsigned i = threadIdx.x; i < BRO_DICT_IMAGE_BYTES / 16u; i += WARPS * 32) ((uint4*)s_dict)[i] = ((const uint4*)p.dict)[i];
    __syncthreads();
#endif
    const unsigned warp = threadIdx.x / BRO_W, lane = bro_lane();     // "warp" = group of BRO_W lanes
    const unsigned gwarp = blockIdx.x * (WARPS * BRO_GROUPS_PER_WARP) + warp;
    // retry pass of the two-phase path: only the streams the parse kernel handed over -- unless AUTO's gate sent the
    // whole batch here
    const bool gated = p.gate && p.gate[1];
    const bool retry = p.retry_mode && !gated;
    if (retry && *p.retry_count == 0u) return;
    if (p.fused_role) {
        const bool small = gated || (retry && *p.retry_count <= p.fused_small);
        if ((p.fused_role == 1) != small) return;      // the other build's job
    }
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
        // (the latency build's launch carries 2 KB more shared memory per warp, behind the scratch blocks and the dictionary image)
        d.root10 = MINB == BRO_MIN_BLOCKS_LATENCY && BRO_W == 32u ? (uint16_t*)(bro_smem_raw + BRO_WARP_KERNEL_SMEM_BASE) + 1024u * warp : (uint16_t*)0;
        d.arena = p.arena + (size_t)gwarp * BRO_GROUP_ARENA_U16;
        d.arena_cap = BRO_GROUP_ARENA_U16;
        d.arena_base = 0;
#if BRO_DICT_SMEM
        d.dict = s_dict;
#else
        d.dict = p.dict;
#endif
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


#if !defined(BRO_WARPSIM)
// blocks_per_sm[0]: the throughput build, [1]: the latency build
extern "C" int bro_warp_kernel_occupancy(int* blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(bro_decode_warp_kernel<BRO_WARPS_PER_CTA, BRO_MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BRO_WARP_KERNEL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bro_decode_warp_kernel<BRO_WARPS_PER_CTA, BRO_MIN_BLOCKS_LATENCY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BRO_WARP_KERNEL_SMEM_LATENCY);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[0], bro_decode_warp_kernel<BRO_WARPS_PER_CTA, BRO_MIN_BLOCKS>,
                                                                            BRO_WARPS_PER_CTA * 32, BRO_WARP_KERNEL_SMEM);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[1], bro_decode_warp_kernel<BRO_WARPS_PER_CTA, BRO_MIN_BLOCKS_LATENCY>,
                                                                            BRO_WARPS_PER_CTA * 32, BRO_WARP_KERNEL_SMEM_LATENCY);
    return (int)e;
}

extern "C" int bro_warp_kernel_warps_per_cta() { return BRO_WARPS_PER_CTA * BRO_GROUPS_PER_WARP; }
extern "C" size_t bro_warp_kernel_arena_bytes() { return 2u * (size_t)BRO_GROUP_ARENA_U16; }

extern "C" int bro_warp_kernel_launch(const BroLaunch* p, int grid, int latency, cudaStream_t stream) {
    (void)cudaGetLastError();      // a stale error of another library in this process (it is per thread) is not this launch's
    if (latency) bro_decode_warp_kernel<BRO_WARPS_PER_CTA, BRO_MIN_BLOCKS_LATENCY><<<grid, BRO_WARPS_PER_CTA * 32, BRO_WARP_KERNEL_SMEM_LATENCY, stream>>>(*p);
    else bro_decode_warp_kernel<BRO_WARPS_PER_CTA, BRO_MIN_BLOCKS><<<grid, BRO_WARPS_PER_CTA * 32, BRO_WARP_KERNEL_SMEM, stream>>>(*p);
    return (int)cudaGetLastError();
}
#endif
