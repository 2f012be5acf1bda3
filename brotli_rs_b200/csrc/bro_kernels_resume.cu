// bro_kernels_resume.cu -- the RESUMABLE decode kernel for sm_100a (bro_batch_decode_resume, the device side of the
// streaming reader bro_reader_*): the warp-per-stream decoder of bro_kernels.cu started from, and leaving behind, a
// resume point per stream (BroResume, bro_records.h).
//
// The reference decodes incrementally: `State` and the decoder's fields (src/lib.rs:245-291, 378-394) survive between
// read() calls, input is pulled as needed and at most a window of output is kept (src/ringbuffer/mod.rs).  Here a call
// decodes as many whole meta-blocks as the input and the output slot it is given hold; a resume point is written in
// front of every meta-block header, so a call that runs out of either inside a meta-block is repeated from there.  The
// slot starts with the history (the last min(window, bytes so far) bytes of output), which is what back-references and
// the distance limit (src/lib.rs:1489) need.
//
// A kernel of its own rather than a mode of bro_decode_warp_kernel, so that the batch kernel's code and register
// allocation stay what was tuned and measured.
#if !defined(BRO_WARPSIM)   /* (BRO_WARPSIM: this kernel compiled for the host, 32 lanes as fibers -- CPU test-suite only, bro_warpsim.cpp) */
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "bro_decoder_core.h"
#include "bro_kernels.h"

#define BRO_RESUME_WARPS 8

__global__ void __launch_bounds__(BRO_RESUME_WARPS * 32, 2) bro_decode_resume_kernel(BroLaunch p) {      // 128 registers: a reader's one stream is latency-bound
#if defined(BRO_WARPSIM)
    uint8_t* const bro_smem_raw = ws_dynamic_smem;
#else
    extern __shared__ __align__(16) uint8_t bro_smem_raw[];
#endif
    BroScratch* const scratch = (BroScratch*)bro_smem_raw;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned gwarp = blockIdx.x * BRO_RESUME_WARPS + warp;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(p.counter, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= p.n) break;
        const uint64_t in_b = p.in_off[i], in_e = p.in_off[i + 1];
        const uint64_t out_b = p.out_off[i], out_e = p.out_off[i + 1];
        BroDec d;
        d.sc = &scratch[warp];
        d.root10 = 0;
        d.arena = p.arena + (size_t)gwarp * BRO_ARENA_U16_MAX;
        d.arena_cap = BRO_ARENA_U16_MAX;
        d.arena_base = 0;
        d.dict = p.dict;
        d.out = p.out + out_b;
        const uint64_t cap = out_e - out_b;
        d.cap = cap > BRO_MAX_SLOT ? (uint32_t)BRO_MAX_SLOT : (uint32_t)cap;
        d.quirk_spec = p.quirk_spec;
        int st = BRO_ST_OutputTooSmall;                       // a history longer than the slot cannot be resumed
        d.pos = p.resume[i].pos;
        if (d.pos <= d.cap) st = bro_decode_stream_resume(d, p.resume + i, p.in + in_b, p.in + in_e);
        __syncwarp();
        if (lane == 0) {
            p.status[i] = st;
            p.out_len[i] = d.pos;
        }
    }
}

#if !defined(BRO_WARPSIM)
extern "C" int bro_resume_kernel_warps_per_cta() { return BRO_RESUME_WARPS; }

extern "C" int bro_resume_kernel_launch(const BroLaunch* p, int grid, cudaStream_t stream) {
    (void)cudaGetLastError();
    static bool attr_set = false;      // (per process; the attribute is per function and device, set before every first use on a device is overkill: it is idempotent)
    cudaError_t e = cudaFuncSetAttribute(bro_decode_resume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BRO_RESUME_WARPS * sizeof(BroScratch)));
    if (e != cudaSuccess) return (int)e;
    (void)attr_set;
    bro_decode_resume_kernel<<<grid, BRO_RESUME_WARPS * 32, BRO_RESUME_WARPS * sizeof(BroScratch), stream>>>(*p);
    return (int)cudaGetLastError();
}
#endif
