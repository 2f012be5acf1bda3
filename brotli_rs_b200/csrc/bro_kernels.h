// Launch interface between the C ABI (bro_abi.cu) and the kernels (bro_kernels.cu, bro_kernels_thread.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct BroLaunch {
    const uint8_t* in;
    const uint64_t* in_off;
    uint8_t* out;
    const uint64_t* out_off;
    uint64_t* out_len;
    int32_t* status;
    uint32_t n;
    uint16_t* arena;          // warp kernel: num_warps * BRO_ARENA_U16_MAX; thread kernel: num_threads * BRO_THREAD_ARENA_U16
    const uint8_t* dict;      // 122,784-byte dictionary image in HBM
    uint32_t* counter;        // work queue head, zeroed before every launch
    const uint32_t* order;    // thread kernel: stream indices, largest compressed size first (NULL = identity)
    uint32_t* retry_count;    // thread kernel increments it per ArenaTooSmall stream; the warp kernel in retry mode
                              // decodes exactly the streams whose status is ArenaTooSmall (and exits at once if 0)
    int retry_mode;
    int quirk_spec;
};

// warp-per-stream kernel (bro_kernels.cu)
extern "C" int bro_warp_kernel_occupancy(int* blocks_per_sm);
extern "C" int bro_warp_kernel_warps_per_cta();
extern "C" size_t bro_warp_kernel_arena_bytes();
extern "C" int bro_warp_kernel_launch(const BroLaunch* p, int grid, cudaStream_t stream);

// thread-per-stream kernel and the size-class ordering kernels (bro_kernels_thread.cu)
extern "C" int bro_thread_kernel_occupancy(int* blocks_per_sm);
extern "C" int bro_thread_kernel_block();
extern "C" size_t bro_thread_kernel_arena_bytes();
extern "C" int bro_thread_kernel_launch(const BroLaunch* p, int grid, cudaStream_t stream);
// order[] <- stream indices grouped by compressed-size class, largest first.  scratch: 512 uint32.
extern "C" int bro_order_launch(const uint64_t* in_off, uint32_t n, uint32_t* order, uint32_t* scratch, cudaStream_t stream);
