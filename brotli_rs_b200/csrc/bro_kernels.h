// Launch interface between the C ABI (bro_abi.cu) and the kernels (bro_kernels.cu).
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
    uint16_t* arena;      // num_warps * BRO_ARENA_U16
    const uint8_t* dict;  // 122,784-byte dictionary image in HBM
    uint32_t* counter;    // work queue head, zeroed before every launch
    int quirk_spec;
};

extern "C" int bro_kernel_occupancy(int* blocks_per_sm);
extern "C" int bro_kernel_warps_per_cta();
extern "C" size_t bro_kernel_arena_bytes_per_warp();
extern "C" int bro_kernel_launch(const BroLaunch* p, int grid, cudaStream_t stream);
