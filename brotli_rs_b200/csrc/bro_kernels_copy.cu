// bro_kernels_copy.cu -- PHASE TWO of the two-phase path for sm_100a: the copy kernel (one WARP per stream).
//
// Input: the copy records phase one wrote for a stream (bro_parse.h), in stream order: LZ77 back-references into the
// output produced so far (src/lib.rs:1483-1505; the ring buffer of src/ringbuffer/mod.rs is the linear output slot
// itself) and stored meta-blocks (src/lib.rs:1701-1734).  Literals and dictionary words are already in the slot.
//
// The records of a stream must take effect in order, but most neighbours do not depend on each other.  The warp loads
// 32 records at a time (one per lane, coalesced) and splits them into GROUPS: a maximal run of records none of which
// reads a byte that a record of the same run writes.  Inside a group nothing has to be ordered:
//   * a group of LONG records (phase one cuts every copy into pieces of at most 32 aligned 16-byte vectors) is moved
//     piece by piece by groups of BRO_COPY_GROUP lanes (bro_copy_piece.h): lane t of a group moves the vectors t,
//     t + GROUP, ... and the ragged bytes at both ends, so a warp step moves 32 / GROUP pieces -- with BRO_COPY_PIECES
//     pieces in flight: all loads are issued before the first store;
//   * a group of SHORT records (text-like streams) is executed as ONE segmented copy: the bytes of all its records are
//     cut into units (16-byte vectors on aligned destinations, single bytes for the ragged ends), the units are
//     numbered by a warp prefix sum, and every lane takes every 32nd unit, finding its record by a binary search over
//     the prefix sums with shuffles; the sources are staged through shared memory by cp.async, several rounds deep.
// A record that overlaps its own source (distance < length, a periodic fill) is executed alone by doubling: each pass
// copies the whole pattern laid down so far.
//
// Roofline: HBM.  Algorithmic bytes per stream = output bytes written by records + 16 bytes per record read; sources
// are re-reads of recent output (L2 hits for windows that fit).
#if !defined(BRO_WARPSIM)   /* (BRO_WARPSIM: this kernel compiled for the host, 32 lanes as fibers -- CPU test-suite only, bro_warpsim_copy.cpp) */
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "bro_copy_piece.h"
#include "bro_decoder_core.h"
#include "bro_kernels.h"

// 2 CTAs of 10 warps per SM: the 102-register cap of that shape is what the kernel needs to hold a step's data without spilling
// (92 registers).  Measured on B200 (profiles/r02_kernel_variants.md section 9): 3 x 8 warps at 80 registers (28 bytes of spills,
// the product until then) 9.00 ms on the headline batch, 2 x 10 warps 8.86 ms (far sources, c7: 1.84 -> 1.72 ms); 4 x 7 warps at 72
// registers 10.6 ms, 5 x 6 at 64 registers 11.8 ms -- every spilled register costs more than the warps it buys.
#ifndef BRO_COPY_WARPS
#define BRO_COPY_WARPS 10
#endif
#ifndef BRO_COPY_MIN_BLOCKS
#define BRO_COPY_MIN_BLOCKS 2
#endif
// The second shape, 3 CTAs of 8 warps (80 registers): 20 % more warps per SM for a batch so small that what it costs is the
// number of streams a warp has to take one after the other (a rank's share of a batch under strong scaling: 12,500 streams are
// 3.5 per warp with this shape, 4.2 -- i.e. five rounds instead of four -- with the first).  The host chooses (bro_abi.cu).
#ifndef BRO_COPY_WARPS_SMALL
#define BRO_COPY_WARPS_SMALL 8
#endif
#ifndef BRO_COPY_MIN_BLOCKS_SMALL
#define BRO_COPY_MIN_BLOCKS_SMALL 3
#endif
#ifndef BRO_COPY_PIECES
#define BRO_COPY_PIECES 4     // long records: pieces in flight per warp (their data is held in registers)
#endif
#ifndef BRO_COPY_GROUP
#define BRO_COPY_GROUP 8      // long records: lanes that move one piece (32, 16 or 8): 32 / GROUP pieces per warp step.
                              // Measured on B200 (profiles/r01c_kernel_variants.md): 8 lanes = 9.1 ms on the headline batch,
                              // 16 = 9.7 ms, 32 = 11.3 ms; choosing per group between 8 and 32 spills and loses (12.7 ms)
#endif
#ifndef BRO_COPY_STAGED
#define BRO_COPY_STAGED 0     // 1: long records through shared memory with per-granule cp.async (measured in round 2: 9.9 ms vs
                              // 9.2 ms on the headline batch -- the issue work of 36 copies per piece eats what the staging wins)
#endif
#ifndef BRO_COPY_BULK
#define BRO_COPY_BULK 0       // 1: long records through shared memory with ONE bulk copy (TMA, cp.async.bulk + mbarrier) per
                              // piece, bro_run_pieces_bulk.  Correct (GPU parity suite) but measured slower in round 2: 10.2 ms at
                              // 40 warps per SM against 9.2 ms for the register path -- 83 warp instructions per piece (spin on
                              // the mbarrier, one piece per consume trip) and the generic -> async proxy fence at every group
                              // (a MEMBAR: 5.4 stall cycles per issue); profiles/r02_kernel_variants.md
#endif
#ifndef BRO_COPY_WINDOW
#define BRO_COPY_WINDOW 0     // 1: the sliding window staged in shared memory (bro_run_pieces_win): a warp keeps the last
                              // BRO_WIN_BYTES of its stream's output in a ring on chip, and a long record whose source lies in the
                              // ring reads it there -- the store -> load round trip through L2 between dependent copies, which is
                              // what bounded the register path (profiles/r02_kernel_variants.md section 2), leaves the chain
#endif
#ifndef BRO_COPY_PIN_TID
#define BRO_COPY_PIN_TID 0
#endif
#ifndef BRO_COPY_PF_DIST
#define BRO_COPY_PF_DIST 8192u   // sources at least this far back (and every stored block: it comes from the compressed
                                 // input) are asked of DRAM with one bulk prefetch (TMA: cp.async.bulk.prefetch.L2) when their
                                 // record is fetched, steps before they are loaded; 0xffffffff = never
#endif
#ifndef BRO_COPY_SLOTS
#define BRO_COPY_SLOTS 8      // bulk path: pieces in flight per warp (a slot of BRO_STAGE_SLOT_BYTES each)
#endif
#ifndef BRO_COPY_QUADS
#define BRO_COPY_QUADS 2      // staged path: warp steps (of 32 / GROUP pieces) in flight
#endif
#ifndef BRO_COPY_DEPTH
#define BRO_COPY_DEPTH 4      // rounds of 32 units in flight per warp; 1 KiB of staging per round and warp
#endif

// Ampere-style asynchronous copy global -> shared, 16 bytes, L2 only (the sources were written moments ago by this very
// SM's write-through stores and are not in L1 anyway).  No register holds the data, so a lane can have several in
// flight; cp.async.wait_all makes the lane's own copies visible to itself.
#if !defined(BRO_WARPSIM)   /* (the simulation supplies these three and has no use for prefetches) */
#define BRO_PREFETCH_BULK_L2(addr, bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(addr), "r"(bytes) : "memory")
#define BRO_PREFETCH_L2(ptr) asm volatile("prefetch.global.L2 [%0];" :: "l"(ptr))
__device__ __forceinline__ void bro_cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void bro_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint4 bro_lds128(uint32_t smem_addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
    return v;
}
#endif

// 16 bytes from an arbitrary address: two aligned 16-byte loads and a funnel (the bytes before/after the 16 wanted
// ones lie in the same 16-byte granules as wanted bytes, i.e. inside the same allocation)
__device__ __forceinline__ uint4 bro_ldu16(const uint8_t* p) {
    const uintptr_t a = (uintptr_t)p;
    const uint4* q = (const uint4*)(a & ~(uintptr_t)15);
    const unsigned sh = (unsigned)(a & 15u);
    uint4 A = q[0];
    if (sh == 0u) return A;
    const uint4 B = q[1];
    uint32_t w0 = A.x, w1 = A.y, w2 = A.z, w3 = A.w, w4 = B.x, w5 = B.y, w6 = B.z, w7 = B.w;
    if (sh & 8u) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = w7; }
    if (sh & 4u) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
    const unsigned bs = 8u * (sh & 3u);
    uint4 r;
    r.x = __funnelshift_r(w0, w1, bs); r.y = __funnelshift_r(w1, w2, bs);
    r.z = __funnelshift_r(w2, w3, bs); r.w = __funnelshift_r(w3, w4, bs);
    return r;
}

// One record on its own, by the whole warp: dst[0..n) = src[0..n), no overlap.
__device__ __forceinline__ void bro_warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, unsigned lane) {
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    if (lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const uint32_t nv = n >> 4;
    for (uint32_t v = lane; v < nv; v += 32u) *(uint4*)(dst + 16u * v) = bro_ldu16(src + 16u * v);
    const uint32_t done = nv << 4;
    if (done + lane < n) dst[done + lane] = src[done + lane];
}

// A group of LONG records [j, e) (lane k holds record k: destination offset, geometry word, source address): G lanes
// move one piece, so the warp moves PP = 32 / G pieces per step, with ROUNDS steps in flight (all loads are issued
// before the first store).  Nothing in a group is ordered.
template <int G>
__device__ __forceinline__ void bro_run_pieces(uint8_t* out, uint32_t dst, uint32_t geo, uint32_t sp_lo, uint32_t sp_hi,
                                               unsigned lane, uint32_t j, uint32_t e) {
    constexpr int PP = 32 / G, ROUNDS = (BRO_COPY_PIECES + PP - 1) / PP;
    const uint32_t bl = lane & (uint32_t)(G - 1), sub = lane / (uint32_t)G;
    for (uint32_t k0 = j; k0 < e; k0 += (uint32_t)(PP * ROUNDS)) {
        BroPieceData<G> D[ROUNDS];
        uint32_t m_dst[ROUNDS], m_geo[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            m_geo[r] = 0;
            if (k0 + (uint32_t)(PP * r) >= e) continue;                  // warp-uniform
            const uint32_t k = k0 + (uint32_t)(PP * r) + sub;            // this lane group's piece
            const int ks = (int)(k & 31u);
            uint32_t g = __shfl_sync(0xffffffffu, geo, ks);
            const uintptr_t s0 = (uintptr_t)__shfl_sync(0xffffffffu, sp_lo, ks) |
                                 ((uintptr_t)__shfl_sync(0xffffffffu, sp_hi, ks) << 32);
            m_dst[r] = __shfl_sync(0xffffffffu, dst, ks);
            if (PP > 1 && k >= e) g = 0;                                 // the group ends inside this step
            m_geo[r] = g;
            bro_piece_load<G>(D[r], (const uint8_t*)s0, g, bl);
        }
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            if (k0 + (uint32_t)(PP * r) >= e) continue;                  // warp-uniform
            bro_piece_store<G>(D[r], out + m_dst[r], m_geo[r], bl);
        }
    }
}

// The same group through shared memory (BRO_COPY_STAGED): no register holds a piece's data between its loads and its
// stores, so the granules of BRO_COPY_QUADS warp steps are in flight at once -- step q + QUADS - 1 is issued before step q
// is consumed -- and the kernel needs far fewer registers.  stage: this warp's QUADS * (32 / G) slots.
__device__ __forceinline__ void bro_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bro_cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int G>
__device__ __forceinline__ void bro_run_pieces_staged(uint8_t* out, uint32_t dst, uint32_t geo, uint32_t sp_lo, uint32_t sp_hi,
                                                      unsigned lane, uint32_t j, uint32_t e, uint32_t stage) {
    constexpr int PP = 32 / G, NQ = BRO_COPY_QUADS;
    const uint32_t bl = lane & (uint32_t)(G - 1), sub = lane / (uint32_t)G;
    const uint32_t nq = (e - j + (uint32_t)PP - 1u) / (uint32_t)PP;
    for (uint32_t q = 0; q < nq + (uint32_t)(NQ - 1); q++) {
        if (q < nq) {                                                    // warp-uniform: issue step q
            const uint32_t k = j + q * (uint32_t)PP + sub;
            const int ks = (int)(k & 31u);
            uint32_t g = __shfl_sync(0xffffffffu, geo, ks);
            const uintptr_t s0 = (uintptr_t)__shfl_sync(0xffffffffu, sp_lo, ks) |
                                 ((uintptr_t)__shfl_sync(0xffffffffu, sp_hi, ks) << 32);
            if (k >= e) g = 0;
            bro_piece_issue<G>(stage + ((q % (uint32_t)NQ) * (uint32_t)PP + sub) * BRO_STAGE_SLOT_BYTES, (const uint8_t*)s0, g, bl);
        }
        bro_cp_async_commit();                                           // one group per trip (an empty one behind the last step)
        if (q + 1u < (uint32_t)NQ) continue;                             // warp-uniform: the pipeline is filling
        const uint32_t c = q - (uint32_t)(NQ - 1);                       // consume step c: everything but the NQ - 1 youngest groups has landed
        bro_cp_async_wait_group<NQ - 1>();
        __syncwarp();                                                    // ... for every lane of the warp
        const uint32_t k = j + c * (uint32_t)PP + sub;
        const int ks = (int)(k & 31u);
        uint32_t g = __shfl_sync(0xffffffffu, geo, ks);
        const uint32_t d = __shfl_sync(0xffffffffu, dst, ks);
        if (k >= e) g = 0;
        bro_piece_consume<G>(stage + ((c % (uint32_t)NQ) * (uint32_t)PP + sub) * BRO_STAGE_SLOT_BYTES, out + d, g, bl);
        __syncwarp();                                                    // the slots of step c are free for step c + NQ
    }
}

// ---- the same group with ONE bulk copy per piece (BRO_COPY_BULK, the product) ----
// A piece's source -- one contiguous range of at most 36 aligned 16-byte granules -- is exactly what the TMA unit's 1-D
// bulk copy moves: cp.async.bulk global -> shared, completion counted in bytes on an mbarrier.  The lane that holds a
// record issues the copy of its piece (one instruction instead of 36 per-granule copies or 8 loads per lane), up to
// BRO_COPY_SLOTS pieces of the group are in flight per warp while no register holds any of their data, the warp sleeps
// on the mbarrier, and then all 32 lanes realign one piece after the other out of shared memory into aligned 16-byte
// stores (bro_piece_consume<32>: lane v moves vector v; lanes 0..15 / 16..31 the ragged bytes in front / behind).
__device__ __forceinline__ void bro_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bro_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bro_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n"
                 "BRO_MBAR_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra BRO_MBAR_DONE;\n\t"
                 "bra BRO_MBAR_WAIT;\n"
                 "BRO_MBAR_DONE:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bro_bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

template <int NS>
__device__ __forceinline__ void bro_run_pieces_bulk(uint8_t* out, uint32_t dst, uint32_t geo, uint32_t sp_lo, uint32_t sp_hi,
                                                    unsigned lane, uint32_t j, uint32_t e, uint32_t stage, uint32_t bar, uint32_t& parity) {
    // what the group reads was written by ordinary stores of this warp (earlier groups: the caller's __syncwarp orders
    // them) and is now read by the asynchronous proxy: order the two
    asm volatile("fence.proxy.async;" ::: "memory");
    for (uint32_t k0 = j; k0 < e; k0 += (uint32_t)NS) {
        const uint32_t cn = e - k0 < (uint32_t)NS ? e - k0 : (uint32_t)NS;
        // lane k0 + i holds record k0 + i (a lane's index is its record's index in the batch of 32)
        const bool mine = lane >= k0 && lane < k0 + cn;
        const uintptr_t s0 = (uintptr_t)sp_lo | ((uintptr_t)sp_hi << 32);
        uint32_t bytes = 0;
        if (mine) bytes = ((BRO_GEO_SRC_MIS(geo) + BRO_GEO_HEAD(geo) + 16u * BRO_GEO_NVEC(geo) + BRO_GEO_TAIL(geo) + 15u) >> 4) << 4;
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == k0) bro_mbar_expect_tx(bar, total);
        __syncwarp();
        if (mine) bro_bulk_load(stage + (lane - k0) * BRO_STAGE_SLOT_BYTES, (const void*)(s0 & ~(uintptr_t)15), bytes, bar);
        bro_mbar_wait(bar, parity);
        parity ^= 1u;
#pragma unroll 1
        for (uint32_t i = 0; i < cn; i++) {
            const int ks = (int)(k0 + i);
            const uint32_t g = __shfl_sync(0xffffffffu, geo, ks);
            const uint32_t d = __shfl_sync(0xffffffffu, dst, ks);
            bro_piece_consume<32>(stage + i * BRO_STAGE_SLOT_BYTES, out + d, g, lane);
        }
        __syncwarp();                                                    // every lane has read the slots: they are free for the next pieces
    }
}


// ---- the same group with the sliding window staged in shared memory (BRO_COPY_WINDOW, the product) ----
// The ring `win` (BRO_WIN_BYTES per warp) holds the output positions [max(wlo, whi - BRO_WIN_BYTES), whi) of the stream the
// warp works on.  A step moves up to 32 / G pieces of the group, one per lane group, in two phases with a __syncwarp between
// them: every lane group LOADS its piece -- from the ring when its source lies there (bro_win_holds), else from global memory
// as bro_run_pieces does -- and then STORES it to the output and DEPOSITS the same bytes in the ring, together with the few
// bytes phase one wrote between the record before and this one (literals: fetched from the output when the records were
// loaded, long before they are needed; geo bits 28..30 say how many).  A record with more such bytes in front of it than a
// lane keeps (BRO_GAP_BIG) starts a step of its own and starts the ring afresh at its destination.  Because all loads of a
// step precede its deposits, a source BRO_WIN_BYTES back is still intact when it is read.
// wlo / whi / seen / hit are warp-uniform.  hit / seen: pieces served by the ring / pieces moved (the caller stops depositing
// for a stream whose copies reach further back than the ring).
#define BRO_WIN_NONE 0xffffffffu
template <int G>
__device__ __forceinline__ void bro_run_pieces_win(uint8_t* out, uint32_t out_mis, uint32_t dst, uint32_t end, uint32_t geo,
                                                   uint32_t sp_lo, uint32_t sp_hi, uint32_t gapw, unsigned lane, uint32_t j,
                                                   uint32_t e, uint32_t win, bool dep, uint32_t& wlo, uint32_t& whi, uint32_t& seen, uint32_t& hit) {
    // dep = false: the group is moved without the ring (stored meta-blocks -- their source is the compressed input -- and
    // streams for which the ring was given up); whi == BRO_WIN_NONE then, so that no source is looked for in the ring
    constexpr int PP = 32 / G;
    const uint32_t bl = lane & (uint32_t)(G - 1), sub = lane / (uint32_t)G;
    const uint32_t big = __ballot_sync(0xffffffffu, BRO_GEO_GAP(geo) == BRO_GAP_BIG);
    const uint32_t out_lo = (uint32_t)(uintptr_t)out;
    uint32_t k0 = j;
    while (k0 < e) {
        // the records of this step: k0 .. k0 + m - 1 (a record with a big gap in front of it starts a step)
        uint32_t m = e - k0 < (uint32_t)PP ? e - k0 : (uint32_t)PP;
        const uint32_t bm = (big >> (k0 + 1u)) & ((1u << (m - 1u)) - 1u);
        if (bm) m = (uint32_t)__ffs(bm);
        const bool fresh = dep && (whi == BRO_WIN_NONE || ((big >> k0) & 1u));
        if (fresh) wlo = whi = __shfl_sync(0xffffffffu, dst, (int)k0);        // nothing in front of record k0 is in the ring
        const uint32_t k = k0 + sub;
        const int ks = (int)(k & 31u);
        uint32_t g = __shfl_sync(0xffffffffu, geo, ks);
        const uint32_t s_lo = __shfl_sync(0xffffffffu, sp_lo, ks);
        const uintptr_t s0 = (uintptr_t)s_lo | ((uintptr_t)__shfl_sync(0xffffffffu, sp_hi, ks) << 32);
        const uint32_t d = __shfl_sync(0xffffffffu, dst, ks);
        const uint32_t gw = __shfl_sync(0xffffffffu, gapw, ks);
        if (sub >= m) g = 0;
        uint32_t gn = BRO_GEO_GAP(g);
        if (gn == BRO_GAP_BIG || (fresh && sub == 0u)) gn = 0;                // (in front of wlo: not part of the ring)
        const uint32_t spos = s_lo - out_lo;                                  // output position of the first source byte
        const uint32_t len = BRO_GEO_HEAD(g) + 16u * BRO_GEO_NVEC(g) + BRO_GEO_TAIL(g);
        const bool inwin = dep && g != 0u && bro_win_holds(spos, len, wlo, whi);
        BroPieceData<G> D;
        if (inwin) bro_piece_load_win<G>(D, win, spos + out_mis, g, bl);
        else bro_piece_load<G>(D, (const uint8_t*)s0, g, bl);
        hit += (uint32_t)__popc(__ballot_sync(0xffffffffu, inwin && bl == 0u));
        seen += m;
        __syncwarp();                                                         // every load from the ring precedes the deposits
        if (!dep) bro_piece_store<G>(D, out + d, g, bl);
        else if (g != 0u) {
            bro_piece_store_win<G>(D, out + d, win, d + out_mis, g, bl);
            if (bl < gn) bro_stage_st8(win + ((d + out_mis - gn + bl) & BRO_WIN_MASK), (gw >> (8u * bl)) & 0xffu);
        }
        if (dep) whi = __shfl_sync(0xffffffffu, end, (int)(k0 + m - 1u));
        __syncwarp();                                                         // the deposits are there for the next step's loads
        k0 += m;
    }
}

template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) bro_copy_kernel(BroLaunch p) {
    if (p.gate && p.gate[1]) return;       // AUTO: this batch goes to the fused kernel as a whole
    // (values derived from %tid come out of a shuffle so that ptxas keeps them in registers instead of re-reading the special
    // register and redoing the arithmetic wherever they are used)
#if BRO_COPY_PIN_TID
    const unsigned tid = __shfl_sync(0xffffffffu, threadIdx.x, threadIdx.x & 31u);
#else
    const unsigned tid = threadIdx.x;
#endif
    const unsigned lane = tid & 31u;
    // staging: per warp BRO_COPY_DEPTH rounds x 32 lanes x 32 bytes (the two aligned 16-byte granules that hold a
    // unit's 16 source bytes)
    // (the staged long-record path uses the same bytes as BRO_COPY_QUADS * (32 / GROUP) piece slots; a warp is in one
    // path at a time)
    constexpr uint32_t STAGE_SHORT = BRO_COPY_DEPTH * 32u * 32u;
    // (the window path keeps its ring in the same bytes: a group that goes through the short-record path starts the ring afresh)
    constexpr uint32_t STAGE_LONG = BRO_COPY_BULK ? BRO_COPY_SLOTS * BRO_STAGE_SLOT_BYTES :
                                    BRO_COPY_STAGED ? BRO_COPY_QUADS * (32u / BRO_COPY_GROUP) * BRO_STAGE_SLOT_BYTES :
                                    BRO_COPY_WINDOW ? BRO_WIN_BYTES : 0u;
    constexpr uint32_t STAGE_WARP = STAGE_SHORT > STAGE_LONG ? STAGE_SHORT : STAGE_LONG;
    __shared__ __align__(128) uint8_t stage[WARPS * STAGE_WARP];
#if BRO_COPY_PIN_TID
    const uint32_t stage_base = __shfl_sync(0xffffffffu, (uint32_t)__cvta_generic_to_shared(stage) + (tid >> 5) * STAGE_WARP, 0);
#else
    const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(stage) + (tid >> 5) * STAGE_WARP;
#endif
#if BRO_COPY_BULK
    // one mbarrier per warp: the bulk copies of a step complete on it (phase parity kept in a register)
    __shared__ __align__(8) unsigned long long mbar[WARPS];
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[tid >> 5]);
    uint32_t bar_parity = 0;
    if (lane == 0) {
        bro_mbar_init(bar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#endif
    // Streams come from the parse kernel's completion queue: a warp takes a ticket (one ahead of the stream it works on)
    // and waits until the slot of that ticket holds a stream index -- immediately, when the parse kernel has already
    // ended; with the two kernels side by side this is where the copy kernel follows the parse kernel's progress.  A
    // stream's descriptors (status, record count, offsets) are then fetched by four lanes at once.
    const uint64_t in_off0 = p.in_off[0];
    uint32_t t_next = 0;
    if (lane == 0) t_next = atomicAdd(p.counter, 1u);
    for (;;) {
        const uint32_t ticket = __shfl_sync(0xffffffffu, t_next, 0);
        if (ticket >= p.n) break;
        uint32_t i = 0xffffffffu;
        if (lane == 0) {
            t_next = atomicAdd(p.counter, 1u);
            const long long t0 = clock64();
            while ((i = ((volatile uint32_t*)p.done_q)[ticket]) == 0xffffffffu) {
                __nanosleep(500);
                if (clock64() - t0 > p.watchdog) { atomicExch(p.fault, 1u); break; }     // watchdog: never spin forever
            }
            __threadfence();
        }
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= p.n) break;                                  // watchdog fired: give up; the fault flag fails the whole batch (bro_abi.cu)
        uint64_t meta = 0;
        if (lane == 0) meta = (uint64_t)(uint32_t)p.status[i];
        else if (lane == 1) meta = p.nrec[i];
        else if (lane == 2) meta = p.in_off[i];
        else if (lane == 3) meta = p.out_off[i];
        if (__shfl_sync(0xffffffffu, (uint32_t)meta, 0) != (uint32_t)BRO_ST_OK) continue;   // bytes of a failed stream are not part of the contract
        const uint32_t n = __shfl_sync(0xffffffffu, (uint32_t)meta, 1);
        if (n == 0u) continue;
        const uint64_t in_b = __shfl_sync(0xffffffffu, meta, 2);
        const BroRec* recs = p.rec + (((in_b - in_off0) >> 1) + 32ull * i);      // BRO_REC_BASE
        uint8_t* const out = p.out + __shfl_sync(0xffffffffu, meta, 3);
        const uint8_t* const in = p.in + in_b;
        const uint32_t out_mis = (uint32_t)((uintptr_t)out & 15u);   // destination alignment is that of the address
        unsigned long long moved = 0;                                // bytes this lane's records move (measurement)
#if BRO_COPY_WINDOW
        // the ring of this stream (bro_run_pieces_win): empty; tail_end = where the last record of the batch before ended
        uint32_t wlo = 0, whi = BRO_WIN_NONE, w_seen = 0, w_hit = 0, tail_end = 0;
        bool win_on = true;
#endif
        for (uint32_t b = 0; b < n; b += 32u) {
            const uint32_t cnt = n - b < 32u ? n - b : 32u;
            uint32_t dst = 0, lk = 0, a = 0;
            if (lane < cnt) {
                const uint4 r = __ldg((const uint4*)(recs + b + lane));
                dst = r.x; lk = r.y; a = r.z;
            }
            const uint32_t len = lk & BRO_REC_LEN_MASK, kind = lk >> BRO_REC_KIND_SHIFT;
            moved += len;
#if BRO_COPY_WINDOW
            // what phase one wrote between the record before and this one (the literals of the command, a dictionary word):
            // up to four such bytes are fetched now -- nothing on the way to them depends on this kernel's stores -- and go
            // into the ring with the record
            const uint32_t end = dst + len;
            uint32_t gapcode = 0, gapw = 0;
            if (win_on) {
                uint32_t prev = __shfl_up_sync(0xffffffffu, end, 1);
                if (lane == 0) prev = tail_end;
                const uint32_t gap = dst - prev;
                gapcode = gap <= 4u ? gap : BRO_GAP_BIG;
                if (lane < cnt && gap - 1u < 4u) {
                    uint32_t gb[4];
#pragma unroll
                    for (uint32_t i = 0; i < 4u; i++) gb[i] = i < gap ? out[prev + i] : 0u;
                    gapw = gb[0] | (gb[1] << 8) | (gb[2] << 16) | (gb[3] << 24);
                }
                tail_end = __shfl_sync(0xffffffffu, end, (int)(cnt - 1u));
            }
#endif
            // What bounds this kernel is the chain of memory round trips along a stream (a group's loads follow the stores
            // of the group before it), and a round trip that misses L2 is three times as long.  Output written a few KB ago
            // is still in L2; far sources -- written tens of microseconds ago and evicted by the 26 GB that followed -- and
            // stored blocks are not.  They are known from the records long before they are needed: one bulk prefetch per
            // record into L2 now, and the loads of the steps that follow find them there (round 1: 28 % of the source
            // sectors came from DRAM, on the critical path).
            if (lane < cnt && (kind == BRO_REC_STORED || (a >= BRO_COPY_PF_DIST && a >= len))) {
                const uintptr_t s0 = (uintptr_t)(kind == BRO_REC_STORED ? in + a : (const uint8_t*)out + (dst - a));
                const uint32_t span = len < 1024u ? len : 1024u;
                BRO_PREFETCH_BULK_L2(s0 & ~(uintptr_t)15, (((uint32_t)s0 & 15u) + span + 15u) & ~15u);
            }
            if (b + 32u + lane < n) BRO_PREFETCH_L2(recs + b + 32u + lane);     // the next records
            uint32_t j = 0;
            while (j < cnt) {
                // the group [j, e): no record reads what a record of the group writes
                const uint32_t dst_j = __shfl_sync(0xffffffffu, dst, j);
                const bool indep = lane >= j && lane < cnt && (kind == BRO_REC_STORED || (dst - a) + len <= dst_j);
                const uint32_t fail = ~__ballot_sync(0xffffffffu, indep) & (0xffffffffu << j);
                const uint32_t e = fail ? (uint32_t)__ffs(fail) - 1u : 32u;
                __syncwarp();      // stores of earlier groups are visible to the loads below
                if (e == j) {
                    // a record that overlaps its own source: periodic fill by doubling
                    const uint32_t dj = dst_j, lj = __shfl_sync(0xffffffffu, len, j), aj = __shfl_sync(0xffffffffu, a, j);
                    uint32_t done = 0;
                    while (done < lj) {
                        uint32_t m = done + aj;                    // bytes of pattern laid down so far
                        if (m > lj - done) m = lj - done;
                        bro_warp_copy(out + dj + done, out + dj - aj, m, lane);
                        done += m;
                        __syncwarp();
                    }
                    j += 1u;
#if BRO_COPY_WINDOW
                    whi = BRO_WIN_NONE;
#endif
                    continue;
                }
                // units of my record: head bytes up to the next 16-byte boundary, 16-byte vectors, tail bytes
                const bool mine = lane >= j && lane < e;
                uint32_t head = (16u - ((dst + out_mis) & 15u)) & 15u;
                if (head > len) head = len;
                const uint32_t nvec = (len - head) >> 4, tail = (len - head) & 15u;
                // LONG records (phase one cuts copies into pieces of <= 32 vectors): one piece per warp step -- lane t
                // moves vector t, lanes 0..15 the ragged bytes in front, lanes 16..31 those behind -- with
                // BRO_COPY_PIECES pieces in flight (all loads are issued before the first store) and no search.
                const uint32_t gsum = __reduce_add_sync(0xffffffffu, mine ? len : 0u);
                if (gsum >= 192u * (e - j) && !__any_sync(0xffffffffu, mine && nvec > 32u)) {
                    // per record (lane-local, broadcast per piece): source address and geometry
                    const uint8_t* sp = kind == BRO_REC_STORED ? in + a : (const uint8_t*)out + (dst - a);
                    const uint32_t sp_lo = (uint32_t)(uintptr_t)sp, sp_hi = (uint32_t)((uintptr_t)sp >> 32);
#if BRO_COPY_WINDOW
                    const uint32_t geo = bro_piece_geo(dst + out_mis, sp_lo, len) | (gapcode << 28);
                    const bool dep = win_on && !__any_sync(0xffffffffu, mine && kind == BRO_REC_STORED);
                    if (!dep) whi = BRO_WIN_NONE;
                    bro_run_pieces_win<BRO_COPY_GROUP>(out, out_mis, dst, end, geo, sp_lo, sp_hi, gapw, lane, j, e, stage_base, dep,
                                                       wlo, whi, w_seen, w_hit);
                    // a stream whose copies reach further back than the ring: stop depositing
                    if (w_seen >= 64u && 4u * w_hit < w_seen) win_on = false;
#else
                    const uint32_t geo = bro_piece_geo(dst + out_mis, sp_lo, len);
#if BRO_COPY_BULK
                    bro_run_pieces_bulk<BRO_COPY_SLOTS>(out, dst, geo, sp_lo, sp_hi, lane, j, e, stage_base, bar, bar_parity);
#elif BRO_COPY_STAGED
                    bro_run_pieces_staged<BRO_COPY_GROUP>(out, dst, geo, sp_lo, sp_hi, lane, j, e, stage_base);
#else
                    bro_run_pieces<BRO_COPY_GROUP>(out, dst, geo, sp_lo, sp_hi, lane, j, e);
#endif
#endif
                    j = e;
                    continue;
                }
                // SHORT records: one segmented copy over all units of the group
#if BRO_COPY_WINDOW
                whi = BRO_WIN_NONE;       // (its staging slots are the ring's bytes)
#endif
                const uint32_t units = mine ? head + nvec + tail : 0u;
                uint32_t incl = units;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)lane >= o) incl += v;
                }
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                // BRO_COPY_DEPTH rounds of 32 units are in flight at a time: ISSUE (find the unit's record, start the
                // asynchronous copy of its 32 source bytes into this lane's staging slot), wait once, then STORE
                // (realign from the slot, one aligned 16-byte store).  The load latency is paid once per
                // BRO_COPY_DEPTH rounds instead of once per round.
                for (uint32_t u0 = 0; u0 < total; u0 += 32u * BRO_COPY_DEPTH) {
                    uint32_t m_dp[BRO_COPY_DEPTH], m_info[BRO_COPY_DEPTH];   // destination offset; 0 none | 1 byte | 2 vector, source shift << 8
                    uint32_t m_byte[BRO_COPY_DEPTH];                         // the byte of a byte unit (not touched before the wait)
#pragma unroll
                    for (int r = 0; r < BRO_COPY_DEPTH; r++) {
                        const uint32_t u = u0 + 32u * r + lane;
                        m_info[r] = 0; m_dp[r] = 0; m_byte[r] = 0;
                        if (u0 + 32u * r >= total) continue;      // warp-uniform
                        // smallest k with incl[k] > u
                        uint32_t k = 0;
#pragma unroll
                        for (int sft = 16; sft >= 1; sft >>= 1) {
                            const uint32_t v = __shfl_sync(0xffffffffu, incl, (int)(k + sft - 1u));
                            if (v <= u) k += (uint32_t)sft;
                        }
                        const int ks = (int)(k & 31u);
                        const uint32_t r_dst = __shfl_sync(0xffffffffu, dst, ks), r_lk = __shfl_sync(0xffffffffu, lk, ks);
                        const uint32_t r_a = __shfl_sync(0xffffffffu, a, ks), r_end = __shfl_sync(0xffffffffu, incl, ks);
                        if (u < total) {
                            const uint32_t r_len = r_lk & BRO_REC_LEN_MASK, r_kind = r_lk >> BRO_REC_KIND_SHIFT;
                            uint32_t r_head = (16u - ((r_dst + out_mis) & 15u)) & 15u;
                            if (r_head > r_len) r_head = r_len;
                            const uint32_t r_nvec = (r_len - r_head) >> 4;
                            const uint32_t r_units = r_head + r_nvec + ((r_len - r_head) & 15u);
                            const uint32_t ul = u - (r_end - r_units);         // unit index inside the record
                            uint32_t off;                                      // byte offset inside the record
                            bool vec = false;
                            if (ul < r_head) off = ul;
                            else if (ul < r_head + r_nvec) { off = r_head + 16u * (ul - r_head); vec = true; }
                            else off = r_head + 16u * r_nvec + (ul - r_head - r_nvec);
                            const uint8_t* src = r_kind == BRO_REC_STORED ? in + r_a + off : out + (r_dst - r_a) + off;
                            m_dp[r] = r_dst + off;
                            if (vec) {
                                const uintptr_t sa = (uintptr_t)src;
                                const uint32_t sh = (uint32_t)(sa & 15u);
                                const uint32_t slot = stage_base + (uint32_t)(r * 32 + (int)lane) * 32u;
                                bro_cp_async16(slot, (const void*)(sa & ~(uintptr_t)15));
                                if (sh) bro_cp_async16(slot + 16u, (const void*)((sa & ~(uintptr_t)15) + 16u));
                                m_info[r] = 2u | (sh << 8);
                            } else { m_info[r] = 1u; m_byte[r] = *src; }
                        }
                    }
                    bro_cp_async_wait_all();
#pragma unroll
                    for (int r = 0; r < BRO_COPY_DEPTH; r++) {
                        const uint32_t info = m_info[r];
                        if ((info & 3u) == 2u) {
                            const uint32_t slot = stage_base + (uint32_t)(r * 32 + (int)lane) * 32u;
                            const uint32_t sh = info >> 8;
                            uint4 A = bro_lds128(slot);
                            if (sh) {
                                const uint4 B = bro_lds128(slot + 16u);
                                uint32_t w0 = A.x, w1 = A.y, w2 = A.z, w3 = A.w, w4 = B.x, w5 = B.y, w6 = B.z, w7 = B.w;
                                if (sh & 8u) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = w7; }
                                if (sh & 4u) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
                                const unsigned bs = 8u * (sh & 3u);
                                A.x = __funnelshift_r(w0, w1, bs); A.y = __funnelshift_r(w1, w2, bs);
                                A.z = __funnelshift_r(w2, w3, bs); A.w = __funnelshift_r(w3, w4, bs);
                            }
                            *(uint4*)(out + m_dp[r]) = A;
                        } else if (info & 1u) out[m_dp[r]] = (uint8_t)m_byte[r];
                    }
                }
                j = e;
            }
        }
        // what the roofline of this kernel is computed from: record bytes moved and records read (bench.py)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) moved += __shfl_xor_sync(0xffffffffu, moved, o);
        if (lane == 0) {
            atomicAdd(p.copy_stats, moved);
            atomicAdd(p.copy_stats + 1, (unsigned long long)n);
        }
    }
}

#if !defined(BRO_WARPSIM)
// shape 0: the throughput shape, 1: the small-batch shape
extern "C" int bro_copy_kernel_occupancy(int* blocks_per_sm) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[0], bro_copy_kernel<BRO_COPY_WARPS, BRO_COPY_MIN_BLOCKS>,
                                                                  BRO_COPY_WARPS * 32, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[1], bro_copy_kernel<BRO_COPY_WARPS_SMALL, BRO_COPY_MIN_BLOCKS_SMALL>,
                                                                            BRO_COPY_WARPS_SMALL * 32, 0);
    return (int)e;
}
extern "C" int bro_copy_kernel_warps_per_cta(int shape) { return shape ? BRO_COPY_WARPS_SMALL : BRO_COPY_WARPS; }

extern "C" int bro_copy_kernel_launch(const BroLaunch* p, int grid, int shape, cudaStream_t stream) {
    (void)cudaGetLastError();
    if (shape) bro_copy_kernel<BRO_COPY_WARPS_SMALL, BRO_COPY_MIN_BLOCKS_SMALL><<<grid, BRO_COPY_WARPS_SMALL * 32, 0, stream>>>(*p);
    else bro_copy_kernel<BRO_COPY_WARPS, BRO_COPY_MIN_BLOCKS><<<grid, BRO_COPY_WARPS * 32, 0, stream>>>(*p);
    return (int)cudaGetLastError();
}
#endif
