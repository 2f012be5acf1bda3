// bro_copy_piece.h -- the lane-local half of the copy kernel's LONG-record path (bro_kernels_copy.cu): what ONE lane
// loads and stores for the piece its lane group works on.  No warp intrinsics in here, so the very same code is run
// lane by lane on the host by the CPU test-suite (bro_hostsim_copy.cpp; never part of the product library).
//
// A piece is a copy record of at most 32 destination-aligned 16-byte vectors with fewer than 16 ragged bytes in front
// (head) and behind (tail); phase one cuts every LZ77 back-reference (src/lib.rs:1483-1505) and every stored
// meta-block (src/lib.rs:1701-1734) into such pieces.  G lanes (32, 16 or 8: a warp, half or quarter of one) move one
// piece: lane bl of the group takes the vectors bl, bl + G, ... and the ragged byte slots bl, bl + G, ... (slots 0..15 are
// the head bytes, 16..31 the tail bytes).  Loading and storing are separate calls so that the caller can issue the
// loads of several pieces before the first store.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BRO_PIECE_FN __device__ __forceinline__
typedef uint4 bro_v16;
BRO_PIECE_FN uint32_t bro_piece_funnel(uint32_t lo, uint32_t hi, unsigned s) { return __funnelshift_r(lo, hi, s); }
#else
#define BRO_PIECE_FN static inline
struct alignas(16) bro_v16 { uint32_t x, y, z, w; };
BRO_PIECE_FN uint32_t bro_piece_funnel(uint32_t lo, uint32_t hi, unsigned s) {
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
#endif

// Geometry word of a piece (computed once per record by the lane that holds the record, broadcast per piece):
// head | vectors << 8 | tail << 16 | (source address of the first vector & 15) << 24 | 1 << 31.
// dst_addr_lo / src_addr_lo: low bits of the addresses of the record's first destination / source byte.
BRO_PIECE_FN uint32_t bro_piece_geo(uint32_t dst_addr_lo, uint32_t src_addr_lo, uint32_t len) {
    uint32_t head = (16u - (dst_addr_lo & 15u)) & 15u;
    if (head > len) head = len;
    const uint32_t nvec = (len - head) >> 4, tail = (len - head) & 15u;
    return head | (nvec << 8) | (tail << 16) | (((src_addr_lo + head) & 15u) << 24) | 0x80000000u;
}
#define BRO_GEO_HEAD(g) ((g) & 0xffu)
#define BRO_GEO_NVEC(g) (((g) >> 8) & 0xffu)
#define BRO_GEO_TAIL(g) (((g) >> 16) & 0xffu)
#define BRO_GEO_SHIFT(g) (((g) >> 24) & 15u)

// 16 bytes that start 4 * WS + bs / 8 bytes into the 32 bytes A | B (WS = 0..3 whole words, bs = 0, 8, 16 or 24 bits)
template <int WS>
BRO_PIECE_FN bro_v16 bro_funnel16(const bro_v16& A, const bro_v16& B, unsigned bs) {
    const uint32_t w[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
    bro_v16 r;
    r.x = bro_piece_funnel(w[WS], w[WS + 1], bs); r.y = bro_piece_funnel(w[WS + 1], w[WS + 2], bs);
    r.z = bro_piece_funnel(w[WS + 2], w[WS + 3], bs); r.w = bro_piece_funnel(w[WS + 3], w[WS + 4], bs);
    return r;
}

// What a lane holds of one piece between load and store
template <int G>
struct BroPieceData {
    bro_v16 A[32 / G], B[32 / G];      // the two aligned 16-byte granules that hold vector bl + G * i of the source
    uint32_t rb[32 / G];               // ragged byte slot bl + G * i
};

// s0: address of the piece's first source byte; g: its geometry word (0 = this lane group has no piece: nothing is
// touched); bl: lane index inside the group.
template <int G>
BRO_PIECE_FN void bro_piece_load(BroPieceData<G>& d, const uint8_t* s0, uint32_t g, uint32_t bl) {
    const uint32_t head = BRO_GEO_HEAD(g), nvec = BRO_GEO_NVEC(g), tail = BRO_GEO_TAIL(g);
#pragma unroll
    for (int i = 0; i < 32 / G; i++) {
        const uint32_t slot = bl + (uint32_t)(G * i);
        const bool back = slot >= 16u;                                   // compile-time for G <= 16
        const uint32_t b = slot & 15u;
        if (b < (back ? tail : head)) d.rb[i] = s0[(back ? head + 16u * nvec : 0u) + b];
    }
    // the aligned granule that holds the first source byte of this lane's first vector; the others follow at G * 16 bytes
    const bro_v16* q = (const bro_v16*)(((uintptr_t)s0 + head) & ~(uintptr_t)15) + bl;
#pragma unroll
    for (int i = 0; i < 32 / G; i++) {
        if (bl + (uint32_t)(G * i) < nvec) {
            d.A[i] = q[G * i];
            if (g & 0x0f000000u) d.B[i] = q[G * i + 1];   // never read past the granule of the piece's last source byte
        }
    }
}

// dp: address of the piece's first destination byte
template <int G>
BRO_PIECE_FN void bro_piece_store(const BroPieceData<G>& d, uint8_t* dp, uint32_t g, uint32_t bl) {
    const uint32_t head = BRO_GEO_HEAD(g), nvec = BRO_GEO_NVEC(g), tail = BRO_GEO_TAIL(g), sh = BRO_GEO_SHIFT(g);
#pragma unroll
    for (int i = 0; i < 32 / G; i++) {
        const uint32_t slot = bl + (uint32_t)(G * i);
        const bool back = slot >= 16u;
        const uint32_t b = slot & 15u;
        if (b < (back ? tail : head)) dp[(back ? head + 16u * nvec : 0u) + b] = (uint8_t)d.rb[i];
    }
    // The source shift is the same for every vector of the piece (all lanes of the group): branch once on its whole
    // words, so that each case is four funnel shifts straight from the right registers.
    bro_v16* const dv = (bro_v16*)(dp + head) + bl;
    const unsigned bs = 8u * (sh & 3u);
#define BRO_PIECE_VECTORS(EXPR)                                             \
    _Pragma("unroll") for (int i = 0; i < 32 / G; i++)                      \
        if (bl + (uint32_t)(G * i) < nvec) dv[G * i] = (EXPR);
    if (sh == 0u) { BRO_PIECE_VECTORS(d.A[i]) }
    else switch (sh >> 2) {
        case 0: BRO_PIECE_VECTORS(bro_funnel16<0>(d.A[i], d.B[i], bs)) break;
        case 1: BRO_PIECE_VECTORS(bro_funnel16<1>(d.A[i], d.B[i], bs)) break;
        case 2: BRO_PIECE_VECTORS(bro_funnel16<2>(d.A[i], d.B[i], bs)) break;
        default: BRO_PIECE_VECTORS(bro_funnel16<3>(d.A[i], d.B[i], bs)) break;
    }
#undef BRO_PIECE_VECTORS
}

// ------------------------------------------------------------------------------------------------------
// The STAGED form of the same piece (BRO_COPY_STAGED; not the shipped configuration -- see profiles/r01c_kernel_variants.md):
// the source of a piece goes through a slot of shared memory instead of registers.  bro_piece_issue starts the copy of
// every aligned 16-byte granule that holds a source byte of the piece into the slot (cp.async on the device: no
// register holds the data, so a lane can have the granules of several pieces in flight); bro_piece_consume, once they
// have landed, realigns from the slot and stores.  A slot holds BRO_STAGE_SLOT_BYTES: (15 + 15 + 512 + 15 + 15) / 16 = 35
// granules at most, rounded to 36.
// ------------------------------------------------------------------------------------------------------
#define BRO_STAGE_SLOT_BYTES 576u

#if defined(__CUDACC__)
typedef uint32_t bro_stage_ptr;                        // shared-memory address
BRO_PIECE_FN void bro_stage_fetch16(bro_stage_ptr dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
BRO_PIECE_FN bro_v16 bro_stage_ld16(bro_stage_ptr a) {
    bro_v16 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
BRO_PIECE_FN uint32_t bro_stage_ld8(bro_stage_ptr a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
#else
typedef uint8_t* bro_stage_ptr;
BRO_PIECE_FN void bro_stage_fetch16(bro_stage_ptr dst, const void* src) { *(bro_v16*)dst = *(const bro_v16*)src; }
BRO_PIECE_FN bro_v16 bro_stage_ld16(bro_stage_ptr a) { return *(const bro_v16*)a; }
BRO_PIECE_FN uint32_t bro_stage_ld8(bro_stage_ptr a) { return *a; }
#endif

// (address of the piece's first source byte) & 15, from the geometry word: sh = (s0 + head) & 15
#define BRO_GEO_SRC_MIS(g) ((BRO_GEO_SHIFT(g) - BRO_GEO_HEAD(g)) & 15u)

// Start fetching the granules of the piece into `slot` (16-byte aligned): lane bl of the group takes granules bl, bl + G, ...
// Every granule fetched holds at least one source byte of the piece, i.e. lies inside the allocation of the source.
template <int G>
BRO_PIECE_FN void bro_piece_issue(bro_stage_ptr slot, const uint8_t* s0, uint32_t g, uint32_t bl) {
    if (g == 0u) return;
    const uint32_t len = BRO_GEO_HEAD(g) + 16u * BRO_GEO_NVEC(g) + BRO_GEO_TAIL(g);
    const uint32_t granules = (BRO_GEO_SRC_MIS(g) + len + 15u) >> 4;
    const uint8_t* gb = (const uint8_t*)((uintptr_t)s0 & ~(uintptr_t)15);
#pragma unroll
    for (int i = 0; i < (36 + G - 1) / G; i++) {
        const uint32_t k = bl + (uint32_t)(G * i);
        if (k < granules) bro_stage_fetch16(slot + 16u * k, gb + 16u * k);
    }
}

// The slot holds the piece's source bytes from offset BRO_GEO_SRC_MIS(g) on.  dp: address of the first destination byte.
template <int G>
BRO_PIECE_FN void bro_piece_consume(bro_stage_ptr slot, uint8_t* dp, uint32_t g, uint32_t bl) {
    const uint32_t head = BRO_GEO_HEAD(g), nvec = BRO_GEO_NVEC(g), tail = BRO_GEO_TAIL(g), sh = BRO_GEO_SHIFT(g);
    const uint32_t mis = BRO_GEO_SRC_MIS(g);
#pragma unroll
    for (int i = 0; i < 32 / G; i++) {
        const uint32_t rs = bl + (uint32_t)(G * i);
        const bool back = rs >= 16u;
        const uint32_t b = rs & 15u;
        if (b < (back ? tail : head)) {
            const uint32_t o = (back ? head + 16u * nvec : 0u) + b;
            dp[o] = (uint8_t)bro_stage_ld8(slot + mis + o);
        }
    }
    // vector v lies `sh` bytes into granule (mis + head) / 16 + v of the slot
    const bro_stage_ptr base = slot + 16u * ((mis + head) >> 4) + 16u * bl;
    bro_v16* const dv = (bro_v16*)(dp + head) + bl;
    const unsigned bs = 8u * (sh & 3u);
#define BRO_PIECE_VECTORS(EXPR)                                                         \
    _Pragma("unroll") for (int i = 0; i < 32 / G; i++)                                  \
        if (bl + (uint32_t)(G * i) < nvec) {                                            \
            const bro_v16 A = bro_stage_ld16(base + 16u * (uint32_t)(G * i));           \
            const bro_v16 B = sh ? bro_stage_ld16(base + 16u * (uint32_t)(G * i) + 16u) : A;   \
            dv[G * i] = (EXPR);                                                         \
        }
    if (sh == 0u) { BRO_PIECE_VECTORS(A) }
    else switch (sh >> 2) {
        case 0: BRO_PIECE_VECTORS(bro_funnel16<0>(A, B, bs)) break;
        case 1: BRO_PIECE_VECTORS(bro_funnel16<1>(A, B, bs)) break;
        case 2: BRO_PIECE_VECTORS(bro_funnel16<2>(A, B, bs)) break;
        default: BRO_PIECE_VECTORS(bro_funnel16<3>(A, B, bs)) break;
    }
#undef BRO_PIECE_VECTORS
}

// ------------------------------------------------------------------------------------------------------
// The WINDOW form of the same piece (BRO_COPY_WINDOW, bro_kernels_copy.cu): the warp keeps the last BRO_WIN_BYTES of the
// stream's output in a ring of shared memory -- the sliding window of src/ringbuffer/mod.rs:8-73, staged on chip.  Every
// piece deposits what it stores (bro_piece_store_win), and a piece whose source lies in the ring takes it from there
// (bro_piece_load_win) instead of HBM / L2: the chain of memory round trips along a stream (a copy reads what the copy before
// it wrote) then runs through shared memory.  Ring coordinates are (output position + (address of the slot & 15)) &
// BRO_WIN_MASK, so that what is 16-byte aligned in the output is 16-byte aligned in the ring and a vector never wraps.
// ------------------------------------------------------------------------------------------------------
#ifndef BRO_WIN_BYTES
#define BRO_WIN_BYTES 4096u
#endif
#define BRO_WIN_MASK (BRO_WIN_BYTES - 1u)
#define BRO_GEO_GAP(g) (((g) >> 28) & 7u)     /* bytes phase one wrote between the record before this one and this one: 0..4, 7 = more */
#define BRO_GAP_BIG 7u

#if defined(__CUDACC__)
BRO_PIECE_FN void bro_stage_st16(bro_stage_ptr a, const bro_v16& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
BRO_PIECE_FN void bro_stage_st8(bro_stage_ptr a, uint32_t b) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(b) : "memory"); }
#else
BRO_PIECE_FN void bro_stage_st16(bro_stage_ptr a, const bro_v16& v) { *(bro_v16*)a = v; }
BRO_PIECE_FN void bro_stage_st8(bro_stage_ptr a, uint32_t b) { *a = (uint8_t)b; }
#endif

// so: ring coordinate (not yet masked) of the piece's first source byte.  Same contents of d as bro_piece_load leaves.
template <int G>
BRO_PIECE_FN void bro_piece_load_win(BroPieceData<G>& d, bro_stage_ptr w, uint32_t so, uint32_t g, uint32_t bl) {
    const uint32_t head = BRO_GEO_HEAD(g), nvec = BRO_GEO_NVEC(g), tail = BRO_GEO_TAIL(g);
#pragma unroll
    for (int i = 0; i < 32 / G; i++) {
        const uint32_t slot = bl + (uint32_t)(G * i);
        const bool back = slot >= 16u;
        const uint32_t b = slot & 15u;
        if (b < (back ? tail : head)) d.rb[i] = bro_stage_ld8(w + ((so + (back ? head + 16u * nvec : 0u) + b) & BRO_WIN_MASK));
    }
    const uint32_t q = ((so + head) & ~15u) + 16u * bl;       // the granule that holds the first source byte of this lane's first vector
#pragma unroll
    for (int i = 0; i < 32 / G; i++) {
        if (bl + (uint32_t)(G * i) < nvec) {
            d.A[i] = bro_stage_ld16(w + ((q + 16u * (uint32_t)(G * i)) & BRO_WIN_MASK));
            if (g & 0x0f000000u) d.B[i] = bro_stage_ld16(w + ((q + 16u * (uint32_t)(G * i) + 16u) & BRO_WIN_MASK));
        }
    }
}

// bro_piece_store + the same bytes into the ring.  dp: address of the piece's first destination byte; dofs: its ring
// coordinate (not yet masked; (dofs + head) & 15 == 0 as ((uintptr_t)dp + head) & 15 == 0).
template <int G>
BRO_PIECE_FN void bro_piece_store_win(const BroPieceData<G>& d, uint8_t* dp, bro_stage_ptr w, uint32_t dofs, uint32_t g, uint32_t bl) {
    const uint32_t head = BRO_GEO_HEAD(g), nvec = BRO_GEO_NVEC(g), tail = BRO_GEO_TAIL(g), sh = BRO_GEO_SHIFT(g);
#pragma unroll
    for (int i = 0; i < 32 / G; i++) {
        const uint32_t slot = bl + (uint32_t)(G * i);
        const bool back = slot >= 16u;
        const uint32_t b = slot & 15u;
        if (b < (back ? tail : head)) {
            const uint32_t o = (back ? head + 16u * nvec : 0u) + b;
            dp[o] = (uint8_t)d.rb[i];
            bro_stage_st8(w + ((dofs + o) & BRO_WIN_MASK), d.rb[i]);
        }
    }
    bro_v16* const dv = (bro_v16*)(dp + head) + bl;
    const uint32_t wv = dofs + head + 16u * bl;
    const unsigned bs = 8u * (sh & 3u);
#define BRO_PIECE_VECTORS(EXPR)                                                         \
    _Pragma("unroll") for (int i = 0; i < 32 / G; i++)                                  \
        if (bl + (uint32_t)(G * i) < nvec) {                                            \
            const bro_v16 v = (EXPR);                                                   \
            dv[G * i] = v;                                                              \
            bro_stage_st16(w + ((wv + 16u * (uint32_t)(G * i)) & BRO_WIN_MASK), v);     \
        }
    if (sh == 0u) { BRO_PIECE_VECTORS(d.A[i]) }
    else switch (sh >> 2) {
        case 0: BRO_PIECE_VECTORS(bro_funnel16<0>(d.A[i], d.B[i], bs)) break;
        case 1: BRO_PIECE_VECTORS(bro_funnel16<1>(d.A[i], d.B[i], bs)) break;
        case 2: BRO_PIECE_VECTORS(bro_funnel16<2>(d.A[i], d.B[i], bs)) break;
        default: BRO_PIECE_VECTORS(bro_funnel16<3>(d.A[i], d.B[i], bs)) break;
    }
#undef BRO_PIECE_VECTORS
}

// Is the source [spos, spos + len) of a back-reference in the ring?  wlo: first position deposited without a break since
// the ring was last started, whi: the position behind the last deposit (what the ring holds is [max(wlo, whi - BRO_WIN_BYTES), whi)).
BRO_PIECE_FN bool bro_win_holds(uint32_t spos, uint32_t len, uint32_t wlo, uint32_t whi) {
    return spos >= wlo && spos + len <= whi && spos + BRO_WIN_BYTES >= whi;
}
