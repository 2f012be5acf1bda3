// Launch interface between the C ABI (bro_abi.cu) and the kernels (bro_kernels.cu, bro_kernels_parse.cu,
// bro_kernels_copy.cu, bro_kernels_resume.cu).
#pragma once
#if !defined(BRO_WARPSIM)   /* (BRO_WARPSIM: a kernel compiled for the host by the CPU test-suite, bro_warpsim_copy.cpp) */
#include <cuda_runtime.h>
#endif
#include <stddef.h>
#include <stdint.h>

#include "bro_records.h"

struct BroLaunch {
    const uint8_t* in;
    const uint64_t* in_off;
    uint8_t* out;
    const uint64_t* out_off;
    uint64_t* out_len;
    int32_t* status;
    uint32_t n;
    uint16_t* arena;          // warp kernel: num_warps * BRO_ARENA_U16_MAX; parse kernel: num_threads * thread arena
    const uint8_t* dict;      // 122,784-byte dictionary image in HBM
    uint32_t* counter;        // work queue head of the kernel being launched, zeroed before every batch
    const uint32_t* order;    // parse kernel: stream indices by compressed-size class, largest first (NULL = identity)
    uint32_t* retry_count;    // parse kernel: +1 per stream it hands to the fused kernel (status ArenaTooSmall, NeedFused
                              // or RecordsFull); the warp kernel in retry mode decodes exactly those (exits at once if 0)
    int retry_mode;
    int fused_role;           // warp kernel behind the two-phase kernels: BOTH builds are launched and each decides on the device
                              // whether the pass is its job -- 1 = the latency build: a batch the gate found bound by its longest
                              // stream, or at most fused_small streams to retry; 2 = the throughput build: the rest; 0 = unconditional
    uint32_t fused_small;
    int quirk_spec;
    // two-phase path: copy records.  Stream i owns records [rec_base(i), rec_base(i+1)) of `rec`, where
    // rec_base(i) = ((in_off[i] - in_off[0]) >> 1) + 32 * i; a stream whose share ends beyond rec_total, or that needs
    // more, is handed to the fused kernel.
    BroRec* rec;
    uint64_t rec_total;
    uint32_t* nrec;           // n: records written for stream i
    // two-phase path: the parse kernel announces every stream it has finished (whatever its status) in done_q, in
    // completion order; the copy kernel takes tickets from counter[] and waits for its slot to be filled, so that the two
    // kernels can run side by side.  done_q is all-ones before the batch.
    uint32_t* done_q;
    // AUTO mode: [0] longest compressed stream of the batch (bytes, saturated), [1] 1 = the batch is bound by its longest
    // stream: the parse and copy kernels return at once and the fused kernel decodes everything.  NULL = no gating.
    uint32_t* gate;
    int sizing;               // parse kernel: only measure the streams (bro_batch_sizes): out_len = decoded size, nothing written
    uint32_t* done_tail;
    uint16_t* roots;          // parse kernel: per-thread decode tables (bro_parse.h) when they live in HBM / L2
    unsigned long long* copy_stats;   // [0] bytes moved by records, [1] records executed (this batch)
    uint32_t lanes;           // parse kernel: streams a warp holds at a time (32 for a full batch; fewer when the batch is smaller
                              // than the resident lanes, so that it spreads over all SMs and a warp's steps serve fewer, busier lanes)
    uint32_t* fault;          // copy kernel: set to 1 when its watchdog fired (a completion-queue slot was never filled): the
                              // batch is then reported as BRO_ST_CudaError instead of leaving streams marked OK without their copies
    long long watchdog;       // copy kernel: cycles to wait for one completion-queue slot
    BroResume* resume;        // resumable decode (bro_kernels_resume.cu): n resume points, read at the start and rewritten
};


// warp-per-stream kernel (bro_kernels.cu)
extern "C" int bro_warp_kernel_occupancy(int* blocks_per_sm);
extern "C" int bro_warp_kernel_warps_per_cta();
extern "C" size_t bro_warp_kernel_arena_bytes();
extern "C" int bro_warp_kernel_launch(const BroLaunch* p, int grid, int latency, cudaStream_t stream);

// resumable warp-per-stream kernel (bro_kernels_resume.cu); same arenas and grid as the warp kernel
extern "C" int bro_resume_kernel_warps_per_cta();
extern "C" int bro_resume_kernel_launch(const BroLaunch* p, int grid, cudaStream_t stream);

// two-phase path, phase one: the parse kernel (one thread per stream) and the size-class ordering kernels
// (bro_kernels_parse.cu)
extern "C" int bro_parse_kernel_occupancy(int* blocks_per_sm);
extern "C" int bro_parse_kernel_block();
extern "C" size_t bro_parse_kernel_arena_bytes();
extern "C" size_t bro_parse_kernel_roots_bytes();
extern "C" int bro_parse_kernel_launch(const BroLaunch* p, int grid, int threads, cudaStream_t stream);   // threads: 0 = the full block
// bro_batch_sizes: turn the internal hand-over statuses into BRO_ST_SizeUnknown
extern "C" int bro_sizes_finish_launch(int32_t* status, uint32_t n, cudaStream_t stream);
// two-phase path, phase two: the copy kernel (one warp per stream) (bro_kernels_copy.cu)
// two shapes of the same kernel (0: 2 CTAs x 10 warps per SM, no spills -- batches that fill the GPU; 1: 3 x 8 warps -- small
// batches, where the streams a warp takes one after the other are what the kernel costs); blocks_per_sm[2]
extern "C" int bro_copy_kernel_occupancy(int* blocks_per_sm);
extern "C" int bro_copy_kernel_warps_per_cta(int shape);
extern "C" int bro_copy_kernel_launch(const BroLaunch* p, int grid, int shape, cudaStream_t stream);
// order[] <- stream indices grouped by compressed-size class, largest first.  scratch: 512 uint32.
extern "C" int bro_order_launch(const uint64_t* in_off, uint32_t n, uint32_t* order, uint32_t* scratch, uint32_t* gate, cudaStream_t stream);
