// bro_kernels_parse.cu -- PHASE ONE of the two-phase path for sm_100a: the parse kernel (one THREAD per stream) and
// the size-class ordering kernels that feed it.
//
// Every lane of a warp owns a different stream and runs the lockstep rounds of bro_parse.h (insert&copy symbol, a few
// literals, distance code, copy), so 32 streams advance per warp instruction on the expensive part of the decode.
// Literals and dictionary words go straight into the output slots; LZ77 back-references and stored meta-blocks become
// BroRec records for the copy kernel (bro_kernels_copy.cu).  In sizing mode (bro_batch_sizes) nothing is written at all:
// the kernel only reports every stream's decoded size.
//
// Meta-block headers (prefix codes, context maps: long structured code, bro_decoder_core.h with a 1-lane "warp") are
// entered by the lanes of a warp TOGETHER: a lane that reaches a meta-block boundary waits (up to BRO_PARSE_PATIENCE
// rounds of the others) until every lane is at a boundary, and lanes whose stream ended pull their next stream at the
// same moment.  Streams are handed out by compressed-size class, so the lanes of a warp hold similar streams.
#if !defined(BRO_WARPSIM)   /* (BRO_WARPSIM: this kernel compiled for the host, 32 lanes as fibers -- CPU test-suite only, bro_warpsim_parse.cpp) */
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#define BRO_THREAD_MODE 1
#define BRO_PARSE 1
#include "bro_parse.h"
#include "bro_kernels.h"

// One CTA per SM: BRO_PARSE_BLOCK threads, each with its lane-interleaved block of BRO_TL_BYTES of shared memory
// (384 x 600 B = 225 KB of the SM's 227 KB).
#ifndef BRO_PARSE_BLOCK
#define BRO_PARSE_BLOCK 384
#endif
#define BRO_PARSE_MIN_BLOCKS 1
#define BRO_PARSE_SMEM (BRO_PARSE_BLOCK * BRO_TL_BYTES)
#ifndef BRO_PARSE_PIN_TID
#define BRO_PARSE_PIN_TID 1
#endif
#ifndef BRO_PARSE_PATIENCE
#define BRO_PARSE_PATIENCE 1024u
#endif
// the stride is an odd number of 128-byte lines, so that the threads' root tables do not pile into a few cache sets
#define BRO_THREAD_ARENA_STRIDE_U16 (BRO_THREAD_ARENA_U16 + 704u)

__global__ void __launch_bounds__(BRO_PARSE_BLOCK, BRO_PARSE_MIN_BLOCKS) bro_parse_kernel(BroLaunch p) {
    if (p.gate && p.gate[1]) return;       // AUTO: this batch goes to the fused kernel as a whole
#if BRO_PARSE_PIN_TID
    // ptxas does not keep values derived from %tid in registers: it re-reads the special register (S2R, tens of cycles) and
    // redoes the arithmetic wherever they are used -- in front of every shared-memory look-up of the decode loop.  A value that
    // comes out of a shuffle cannot be recomputed, so it stays in its register.
    const unsigned tid = __shfl_sync(0xffffffffu, threadIdx.x, threadIdx.x & 31u);
#else
    const unsigned tid = threadIdx.x;
#endif
    const unsigned t = blockIdx.x * BRO_PARSE_BLOCK + tid;
    const unsigned lane = tid & 31u;
    uint16_t* const arena = p.arena + (size_t)t * BRO_THREAD_ARENA_STRIDE_U16;
    BroDec d;
    BroParse ps;
    BroMbInfo mb;
    // shared memory: per warp 32 lane-interleaved blocks (the table reader's scratch while a header is read, the decode
    // tables of the current block types inside a meta-block: bro_decoder_core.h, bro_parse.h), per CTA the
    // insert/copy length table.  Nothing of a thread's working set is in local memory.
#if defined(BRO_WARPSIM)
    uint8_t* const s_blocks = ws_dynamic_smem;
#else
    extern __shared__ __align__(16) uint8_t s_blocks[];
#endif
    __shared__ uint32_t s_ic[48];
    if (tid < 48u) bro_ic_compact_entry(tid, s_ic[tid]);
    __syncthreads();
    uint32_t stream = 0;
    uint32_t waited = 0;          // warp-uniform: trips since the first lane reached a boundary
    bool exhausted = false;       // warp-uniform: the queue is empty
    bool warp_imm = false;        // warp-uniform: some lane decodes its stream in immediate mode (only a header changes that)
    ps.kind = BRO_K_DONE; ps.st = -1;   // st < 0: no stream to report
    {
        BroTl tl;
        tl.base = (uint32_t)__cvta_generic_to_shared(s_blocks) + (tid >> 5) * (32u * BRO_TL_BYTES) + 4u * lane;
#if BRO_PARSE_PIN_TID
        tl.base = __shfl_sync(0xffffffffu, tl.base, lane);      // (the same: the block's address, pinned)
#endif
        bro_scratch_bind(d.scv, tl);
        d.in.ring = tl;
    }
    d.ic = s_ic;
    d.arena = arena;
    d.arena_cap = BRO_THREAD_ARENA_U16;
    d.arena_base = 0;
    d.dict = p.dict;
    d.quirk_spec = p.quirk_spec;
    d.sizing = p.sizing ? 1u : 0u;
    for (;;) {
        const uint32_t at_boundary = __ballot_sync(0xffffffffu, ps.kind >= BRO_K_HEADER);
        if (at_boundary != 0u) {
            waited++;
            if (at_boundary == 0xffffffffu || waited > BRO_PARSE_PATIENCE) {
                waited = 0;
                // lanes whose stream ended report it and pull the next one
                if (ps.kind == BRO_K_DONE && ps.st >= 0) {
                    p.status[stream] = ps.st;
                    p.out_len[stream] = d.pos;
                    if (p.nrec) p.nrec[stream] = d.nrec;
                    if (BRO_ST_IS_RETRY(ps.st)) atomicAdd(p.retry_count, 1u);
                    ps.st = -1;
                    // announce the stream to the copy kernel: everything this thread wrote for it (literals, dictionary
                    // words, records, the three values above) is visible before its index appears in the queue
                    __threadfence();
                    if (p.done_q) ((volatile uint32_t*)p.done_q)[atomicAdd(p.done_tail, 1u)] = stream;
                }
                const uint32_t idle = __ballot_sync(0xffffffffu, ps.kind == BRO_K_DONE) & (p.lanes >= 32u ? 0xffffffffu : (1u << p.lanes) - 1u);
                if (idle != 0u && !exhausted) {
                    const int leader = __ffs(idle) - 1;
                    uint32_t base = 0;
                    if ((int)lane == leader) base = atomicAdd(p.counter, (uint32_t)__popc(idle));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (base + (uint32_t)__popc(idle) >= p.n) exhausted = true;
                    if (ps.kind == BRO_K_DONE && ((idle >> lane) & 1u)) {
                        const uint32_t k = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                        if (k < p.n) {
                            stream = p.order ? p.order[k] : k;
                            const uint64_t in_b = p.in_off[stream], in_e = p.in_off[stream + 1];
                            const uint64_t out_b = p.sizing ? 0 : p.out_off[stream], out_e = p.sizing ? 0 : p.out_off[stream + 1];
                            d.out = p.out + out_b;
                            d.out_mis = (uint32_t)((uintptr_t)d.out & 15u);
                            const uint64_t cap = out_e - out_b;
                            d.cap = (cap > BRO_MAX_SLOT || p.sizing) ? (uint32_t)BRO_MAX_SLOT : (uint32_t)cap;
                            d.pos = 0;
                            d.p1 = 0; d.p2 = 0;
                            d.d0 = 4; d.d1 = 11; d.d2 = 15; d.d3 = 16;   // src/lib.rs:407-408
                            // this stream's share of the record arena: one record per 2 compressed bytes + 32
                            const uint64_t rec_b = BRO_REC_BASE(p.in_off, stream);
                            const uint64_t rec_n = BRO_REC_BASE(p.in_off, stream + 1) - rec_b;
                            d.rec = p.rec + rec_b; d.nrec = 0; d.imm = 0; d.in_base = p.in + in_b;
                            d.rec_cap = rec_n > 0x0fffffffull ? 0x0fffffffu : (uint32_t)rec_n;
                            bro_bits_init(d.in, p.in + in_b, p.in + in_e);
                            bro_parse_begin(ps);
                            if (!p.sizing && rec_b + rec_n > p.rec_total) bro_parse_finish(ps, BRO_ST_RecordsFull);
                            if (in_e - in_b >= (1ull << 28)) bro_parse_finish(ps, BRO_ST_NeedFused);   // 32-bit bit counts
                        }
                    }
                }
                if (__all_sync(0xffffffffu, ps.kind == BRO_K_DONE && ps.st < 0)) break;
                if (ps.kind == BRO_K_HEADER) bro_parse_header(d, ps, mb);
                warp_imm = __any_sync(0xffffffffu, ps.kind < BRO_K_HEADER && d.imm != 0u) != 0;
            }
        } else waited = 0;
        if (warp_imm) bro_parse_round<true>(d, ps, mb);
        else bro_parse_round<false>(d, ps, mb);
    }
}

// ---- size-class ordering: 256 classes (8 per power of two of the compressed size), largest class first ----
__device__ __forceinline__ uint32_t bro_size_class(uint64_t len) {
    uint32_t l = len > 0xffffffffull ? 0xffffffffu : (uint32_t)len;
    uint32_t b = l ? 31u - (uint32_t)__clz(l) : 0u;
    uint32_t sub = b >= 3u ? (l >> (b - 3u)) & 7u : 0u;
    return 255u - (b * 8u + sub);
}

// Batches are made of few size classes (replicas, similar streams): the lanes of a warp that hit the same class are
// counted by one atomic of their leader (match_any) instead of serialising on one address.
__global__ void bro_order_hist_kernel(const uint64_t* in_off, uint32_t n, uint32_t* hist, uint32_t* gate) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t len = i < n ? in_off[i + 1] - in_off[i] : 0;
    const uint32_t c = i < n ? bro_size_class(len) : 0xffffffffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, c);
    if (i < n && (peers & ((1u << (threadIdx.x & 31u)) - 1u)) == 0u) atomicAdd(&hist[c], (uint32_t)__popc(peers));
    if (gate) {
        // longest compressed stream of the batch (one atomic per warp)
        const uint32_t m = __reduce_max_sync(0xffffffffu, len > 0xffffffffull ? 0xffffffffu : (uint32_t)len);
        if ((threadIdx.x & 31u) == 0u && m) atomicMax(&gate[0], m);
    }
}

// BRO_GATE_RATIO: A batch whose longest stream, decoded alone, takes longer than the whole batch would take at full
// throughput is bound by that stream on either path, and then the two-phase path only adds the serial tails of its phases
// (a stream alone: ~2 ms per 4 KB on a fused warp, ~3.5 ms through parse + copy): the fused kernel takes the whole batch.
// The ratio is where the two paths were measured to break even on B200 (round 2: 3,000 headline streams 2.0 ms fused /
// 2.6 ms two-phase, 6,000: 4.0 / 2.8, i.e. at ~4,500 streams of equal size).
#define BRO_GATE_RATIO 4500ull

__global__ void bro_order_scan_kernel(const uint32_t* hist, uint32_t* cursor, const uint64_t* in_off, uint32_t n, uint32_t* gate) {
    __shared__ uint32_t s[256];
    uint32_t t = threadIdx.x;
    if (gate && t == 0) gate[1] = ((uint64_t)gate[0] * BRO_GATE_RATIO > in_off[n] - in_off[0]) ? 1u : 0u;
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
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = i < n ? bro_size_class(in_off[i + 1] - in_off[i]) : 0xffffffffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, c);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (i < n && (int)lane == leader) base = atomicAdd(&cursor[c], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n) order[base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = i;
}

#if !defined(BRO_WARPSIM)   /* (the launchers are the device's; the simulation runs the kernels above CTA by CTA, bro_warpsim_parse.cpp) */
extern "C" int bro_order_launch(const uint64_t* in_off, uint32_t n, uint32_t* order, uint32_t* scratch, uint32_t* gate, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, 512 * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return (int)e;
    uint32_t blocks = (n + 255u) / 256u;
    (void)cudaGetLastError();
    bro_order_hist_kernel<<<blocks, 256, 0, stream>>>(in_off, n, scratch, gate);
    bro_order_scan_kernel<<<1, 256, 0, stream>>>(scratch, scratch + 256, in_off, n, gate);
    bro_order_scatter_kernel<<<blocks, 256, 0, stream>>>(in_off, n, scratch + 256, order);
    return (int)cudaGetLastError();
}

#endif

__global__ void bro_sizes_finish_kernel(int32_t* status, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && BRO_ST_IS_RETRY(status[i])) status[i] = BRO_ST_SizeUnknown;
}

#if !defined(BRO_WARPSIM)
extern "C" int bro_sizes_finish_launch(int32_t* status, uint32_t n, cudaStream_t stream) {
    (void)cudaGetLastError();
    bro_sizes_finish_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(status, n);
    return (int)cudaGetLastError();
}

extern "C" int bro_parse_kernel_occupancy(int* blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(bro_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BRO_PARSE_SMEM);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, bro_parse_kernel, BRO_PARSE_BLOCK, BRO_PARSE_SMEM);
}
extern "C" int bro_parse_kernel_block() { return BRO_PARSE_BLOCK; }
extern "C" size_t bro_parse_kernel_arena_bytes() { return 2u * (size_t)BRO_THREAD_ARENA_STRIDE_U16; }
extern "C" size_t bro_parse_kernel_roots_bytes() { return 0; }   // (round 1 kept literal tables in HBM; everything is on chip now)

// threads: a multiple of 32 up to bro_parse_kernel_block() (0 = that).  The kernel is compiled for the full block (its
// registers), but a launch may bring fewer warps per SM: throughput saturates at 10-11 warps per SM and the twelfth costs
// more in contention than it adds (profiles/r02_kernel_variants.md section 10), so the host sizes the block to the batch.
extern "C" int bro_parse_kernel_launch(const BroLaunch* p, int grid, int threads, cudaStream_t stream) {
    (void)cudaGetLastError();
    if (threads <= 0 || threads > BRO_PARSE_BLOCK) threads = BRO_PARSE_BLOCK;
    threads = (threads + 31) & ~31;
    if (threads < 64) threads = 64;                    // (the CTA's first 48 threads fill the insert/copy length table)
    bro_parse_kernel<<<grid, threads, (size_t)threads * BRO_TL_BYTES, stream>>>(*p);
    return (int)cudaGetLastError();
}
#endif
