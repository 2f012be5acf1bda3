// bro_decoder_core.h -- the Brotli stream decoder executed by ONE WARP per stream.
//
// This is the B200 replacement for the whole of the reference's L2+L1 layers:
//   src/lib.rs:412-2170 (stream state machine and every parse_*/decode_* helper), src/bitreader/mod.rs,
//   src/huffman/{mod.rs,tree/mod.rs}, src/ringbuffer/mod.rs, src/transformation/mod.rs.
// It is not a translation: the reference walks heap-array trees one bit at a time inside a 45-state enum;
// here every lane of a warp carries the same (warp-uniform) decoder state, symbols are decoded with an 8-bit
// root table plus a canonical-code search for longer codes, the compressed bytes are fetched 128 B at a time
// by the whole warp and handed out by shuffle, prefix tables are built by the 32 lanes together, and all
// byte movement (LZ77 copies, pattern fills, stored blocks, dictionary words) is lane-parallel.
//
// The file compiles in three modes:
//   * nvcc, sm_100a, default:        BRO_W = 32: ONE WARP PER STREAM.  Every lane carries the same decoder state;
//     table builds and all byte movement are lane-parallel.  Lowest latency per stream, any arena size: used for
//     small batches and as the general fallback (bro_kernels.cu: bro_decode_warp_kernel).
//   * nvcc, sm_100a, BRO_THREAD_MODE: BRO_W = 1: ONE THREAD PER STREAM.  The same code with a 1-lane "warp": 32
//     streams advance per warp instruction, so the serial entropy decode no longer wastes 31/32 of the issue
//     slots.  Used for large batches, with BRO_PARSE: phase one of the two-phase path (bro_parse.h,
//     bro_kernels_parse.cu: bro_parse_kernel), which records copies instead of making them.
//   * BRO_HOSTSIM:                   BRO_W = 1 on the host, plain C++ -- a simulation of the very same code used
//     ONLY by the CPU test-suite (tests/_build/libbro_hostsim.so) to fuzz the decoder logic against the oracle
//     without a GPU.  It is never linked into libbrotli_b200.so and nothing in the product can reach it.
//
// Results are bit-exact with the reference: same bytes, same error class (status = DecompressorError in enum
// order, src/lib.rs:294-319), including the reference's quirks.
#pragma once
#include <stdint.h>

#include "bro_records.h"
#include "bro_status.h"

#if defined(BRO_HOSTSIM)
#define BRO_W 1u
#define BRO_SERIAL 1
#define BRO_FN static inline
#define BRO_MFN inline
#define BRO_COLD static
#define BRO_TABLE_QUAL static const
#elif defined(BRO_THREAD_MODE)
#define BRO_W 1u
#define BRO_SERIAL 1
#define BRO_FN __device__ __forceinline__
#define BRO_MFN __device__ __forceinline__
#define BRO_COLD static __device__ __noinline__
// every lane indexes the format tables with its own stream's symbol: global memory (L1) serves 32 different addresses in
// one access, the constant cache would replay the load once per distinct address
#define BRO_TABLE_QUAL static __device__ const
#else
#if defined(BRO_GROUP_W)
#define BRO_W BRO_GROUP_W   /* lanes that cooperate on one stream: 32 (a warp), 16, 8 or 4 (sub-warp groups) */
#else
#define BRO_W 32u
#endif
#define BRO_FN __device__ __forceinline__
#define BRO_MFN __device__ __forceinline__
#define BRO_COLD static __device__ __noinline__
#define BRO_TABLE_QUAL static __constant__ const
#endif
#include "bro_tables_generated.h"
// the byte-movement routines are real calls in the group modes (one copy of their code, small hot loops)
#if defined(BRO_SERIAL)
#define BRO_COPY_FN BRO_FN
#else
#define BRO_COPY_FN BRO_COLD
#endif

// ------------------------------------------------------------------------------------------------------
// warp primitives
// ------------------------------------------------------------------------------------------------------
#if defined(BRO_SERIAL)
BRO_FN unsigned bro_lane() { return 0; }
BRO_FN uint32_t bro_shfl(uint32_t v, unsigned) { return v; }
BRO_FN uint32_t bro_match_any(uint32_t) { return 1u; }
BRO_FN uint32_t bro_lanemask_lt() { return 0u; }
BRO_FN void bro_syncwarp() {}
#endif
#if defined(BRO_HOSTSIM)
BRO_FN uint32_t bro_brev(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
BRO_FN uint32_t bro_popc(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
BRO_FN uint32_t bro_funnel_r(uint32_t lo, uint32_t hi, unsigned sh) {
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
#else
#if !defined(BRO_SERIAL)
// A "group" is BRO_W consecutive lanes of a warp; all collectives are restricted to the group, so the groups of
// one warp may diverge from each other freely (independent thread scheduling).
BRO_FN unsigned bro_lane() { return threadIdx.x & (BRO_W - 1u); }
BRO_FN unsigned bro_group_shift() { return threadIdx.x & 31u & ~(BRO_W - 1u); }
BRO_FN uint32_t bro_group_mask() { return (BRO_W == 32u ? 0xffffffffu : ((1u << (BRO_W & 31u)) - 1u)) << bro_group_shift(); }
BRO_FN uint32_t bro_shfl(uint32_t v, unsigned src) { return __shfl_sync(bro_group_mask(), v, (int)src, (int)BRO_W); }
BRO_FN uint32_t bro_match_any(uint32_t v) { return __match_any_sync(bro_group_mask(), v) >> bro_group_shift(); }
BRO_FN uint32_t bro_lanemask_lt() { return (1u << bro_lane()) - 1u; }
#if defined(BRO_WARPSIM)
#define bro_syncwarp() bro_ws_syncwarp_at(__LINE__)   /* the simulation counts the barriers by source line and can leave one out */
#else
BRO_FN void bro_syncwarp() { __syncwarp(bro_group_mask()); }
#endif
#endif
BRO_FN uint32_t bro_brev(uint32_t x) { return __brev(x); }
BRO_FN uint32_t bro_popc(uint32_t x) { return (uint32_t)__popc(x); }
BRO_FN uint32_t bro_funnel_r(uint32_t lo, uint32_t hi, unsigned sh) { return __funnelshift_r(lo, hi, sh); }
#endif

// ------------------------------------------------------------------------------------------------------
// memory layout of the per-warp table arena (HBM, L1/L2 resident while a stream is being decoded)
// ------------------------------------------------------------------------------------------------------
// A prefix-code table ("tree record"), in uint16 units (R = 1 << BRO_ROOT_BITS):
//   [0..R)     root: indexed by the next BRO_ROOT_BITS stream bits; entry = symbol | len<<10 (len 1..BRO_ROOT_BITS),
//              0 = no code starts with these bits (hole), 1 = a longer code does (search by length)
//   [R..R+16)  limit[L]: left-justified 15-bit end of all codes of length <= L (canonical order)
//   [R+16..R+32) base[L]: (int16) index into sorted[] of the first code of length L minus its code value
//   [R+32] single-symbol flag (decode consumes zero bits: src/huffman/tree/mod.rs:87-91)
//   [R+33] the single symbol   [R+34] max code length   [R+35] reserved
//   [R+36..]   sorted[]: symbols in canonical order (by length, then by position)
#define BRO_ROOT_BITS 8u
#define BRO_ROOT_SIZE (1u << BRO_ROOT_BITS)
#define BRO_T_LIMIT (BRO_ROOT_SIZE)
#define BRO_T_BASE (BRO_ROOT_SIZE + 16u)
#define BRO_T_SINGLE (BRO_ROOT_SIZE + 32u)
#define BRO_T_SINGLE_SYM (BRO_ROOT_SIZE + 33u)
#define BRO_T_MAXDEPTH (BRO_ROOT_SIZE + 34u)
#define BRO_T_SORTED (BRO_ROOT_SIZE + 36u)
#define BRO_TREE_U16(alphabet) (((BRO_T_SORTED + (alphabet)) + 7u) & ~7u)
// Distance between the literal tables of a meta-block.  The two-phase path looks tables of context-modelled literals up in
// the arena (bro_parse.h): there a table starts on a 128-byte line (the first 64 root entries = the line a stream keeps hot;
// L2 allocates whole lines, so a hot sector in a line of cold ones costs four times its size).
#if defined(BRO_PARSE)
#define BRO_LIT_STRIDE_U16 ((BRO_TREE_U16(BRO_ALPHA_LIT) + 63u) & ~63u)
#else
#define BRO_LIT_STRIDE_U16 BRO_TREE_U16(BRO_ALPHA_LIT)
#endif

// largest output slot one stream may use: positions are 32-bit and one command may add up to 2 * (2^24 + 22594) bytes
#define BRO_MAX_SLOT 0xf0000000ull

#define BRO_MAX_BLTYPES 256u
#define BRO_ALPHA_LIT 256u
#define BRO_ALPHA_CMD 704u
#define BRO_ALPHA_DIST_MAX 520u   // 16 + 120 + (48 << 3)
#define BRO_ALPHA_BTYPE_MAX 258u
#define BRO_ALPHA_BCOUNT 26u
#define BRO_ALPHA_CMAP_MAX 272u   // 16 + 256

// Worst-case arena (uint16 units): three block-type + block-count tables, context modes, both context maps, the
// temporary context-map table, and 256 literal + 256 insert&copy + 256 distance tables.
#define BRO_ARENA_U16_MAX (3u * (BRO_TREE_U16(BRO_ALPHA_BTYPE_MAX) + BRO_TREE_U16(BRO_ALPHA_BCOUNT)) + 128u + 8192u + 512u + \
                           BRO_TREE_U16(BRO_ALPHA_CMAP_MAX) + BRO_MAX_BLTYPES * (BRO_TREE_U16(BRO_ALPHA_LIT) + \
                           BRO_TREE_U16(BRO_ALPHA_CMD) + BRO_TREE_U16(BRO_ALPHA_DIST_MAX)) + 64u)
// Arena of one THREAD in thread-per-stream mode (64 KiB): enough for e.g. 24 literal + 8 insert&copy + 8 distance
// tables; larger meta-blocks fall back to the warp kernel.  The thread's BroScratch sits at its start.
#define BRO_THREAD_ARENA_U16 32768u

// capacities of the general command loop's on-chip tables (bro_stage_hot)
#define BRO_HOT_CMAP_L 1024u    // literal context map: up to 16 block types x 64 contexts
#define BRO_HOT_CMAP_D 256u     // distance context map: up to 64 block types x 4 contexts
#define BRO_HOT_MODES 64u       // context modes of up to 64 literal block types
#define BRO_HOT_LIT_U16 1536u   // narrow roots of the literal codes: 6 codes x 8 bits, 12 x 7, 24 x 6 or 48 x 5

#if defined(BRO_SERIAL)
// ------------------------------------------------------------------------------------------------------
// One decoder per THREAD: its on-chip storage is a block of BRO_TL_BYTES of shared memory, interleaved WORD BY WORD
// with the blocks of the other 31 lanes of its warp: word w of lane l sits at warp_block + (32 * w + l) * 4.  Whatever
// index a lane looks up, it touches bank l and no other lane does -- the 32 unrelated streams of a warp never
// conflict, and a warp access costs one pass (a per-thread contiguous block would put equal indices of all lanes on
// one bank).  Everything a lane touches per symbol lives here or in registers; local memory (1,232 bytes of stack per
// thread in round 1: 470 KB per SM, i.e. an L2 round trip per access) is not used by the parse kernel any more.
// The host simulation addresses a plain buffer with the very same interleave.
// ------------------------------------------------------------------------------------------------------
struct BroTl {
#if defined(BRO_HOSTSIM)
    uint8_t* base;            // this lane's word 0
#else
    uint32_t base;            // shared-window address of this lane's word 0
#endif
};
BRO_FN uint32_t bro_tl_off(uint32_t b) { return ((b & ~3u) << 5) | (b & 3u); }
#if defined(BRO_HOSTSIM)
BRO_FN uint32_t bro_tl_ld8(BroTl t, uint32_t b) { return t.base[bro_tl_off(b)]; }
BRO_FN void bro_tl_st8(BroTl t, uint32_t b, uint32_t v) { t.base[bro_tl_off(b)] = (uint8_t)v; }
BRO_FN uint32_t bro_tl_ld16(BroTl t, uint32_t b) { const uint8_t* p = t.base + bro_tl_off(b); return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
BRO_FN void bro_tl_st16(BroTl t, uint32_t b, uint32_t v) { uint8_t* p = t.base + bro_tl_off(b); p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
BRO_FN uint32_t bro_tl_ld32(BroTl t, uint32_t b) {
    const uint8_t* p = t.base + bro_tl_off(b);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
BRO_FN void bro_tl_st32(BroTl t, uint32_t b, uint32_t v) {
    uint8_t* p = t.base + bro_tl_off(b);
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
#elif defined(BRO_WARPSIM) && defined(BRO_THREAD_MODE)   /* the parse kernel compiled for the host (bro_warpsim_parse.cpp, CPU test-suite): the same window addresses */
BRO_FN uint32_t bro_tl_ld8(BroTl t, uint32_t b) { return *ws_smem_ptr(t.base + bro_tl_off(b)); }
BRO_FN void bro_tl_st8(BroTl t, uint32_t b, uint32_t v) { *ws_smem_ptr(t.base + bro_tl_off(b)) = (uint8_t)v; }
BRO_FN uint32_t bro_tl_ld16(BroTl t, uint32_t b) { return *(const uint16_t*)ws_smem_ptr(t.base + bro_tl_off(b)); }
BRO_FN void bro_tl_st16(BroTl t, uint32_t b, uint32_t v) { *(uint16_t*)ws_smem_ptr(t.base + bro_tl_off(b)) = (uint16_t)v; }
BRO_FN uint32_t bro_tl_ld32(BroTl t, uint32_t b) { return *(const uint32_t*)ws_smem_ptr(t.base + (b << 5)); }
BRO_FN void bro_tl_st32(BroTl t, uint32_t b, uint32_t v) { *(uint32_t*)ws_smem_ptr(t.base + (b << 5)) = v; }
#else
BRO_FN uint32_t bro_tl_ld8(BroTl t, uint32_t b) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(t.base + bro_tl_off(b))); return v; }
BRO_FN void bro_tl_st8(BroTl t, uint32_t b, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(t.base + bro_tl_off(b)), "r"(v) : "memory"); }
BRO_FN uint32_t bro_tl_ld16(BroTl t, uint32_t b) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(t.base + bro_tl_off(b))); return v; }
BRO_FN void bro_tl_st16(BroTl t, uint32_t b, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(t.base + bro_tl_off(b)), "h"((uint16_t)v) : "memory"); }
BRO_FN uint32_t bro_tl_ld32(BroTl t, uint32_t b) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(t.base + (b << 5))); return v; }
BRO_FN void bro_tl_st32(BroTl t, uint32_t b, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(t.base + (b << 5)), "r"(v) : "memory"); }
#endif

// arrays inside the block: sc.cnt[L] reads and sc.cnt[L] = v writes go through the interleave
template <typename T, uint32_t OFF>
struct BroTlArray {
    BroTl t;
    struct Ref {
        BroTl t; uint32_t b;
        BRO_MFN operator T() const {
            return sizeof(T) == 1 ? (T)bro_tl_ld8(t, b) : sizeof(T) == 2 ? (T)(uint16_t)bro_tl_ld16(t, b) : (T)bro_tl_ld32(t, b);
        }
        BRO_MFN const Ref& operator=(T v) const {
            if (sizeof(T) == 1) bro_tl_st8(t, b, (uint32_t)(uint8_t)v);
            else if (sizeof(T) == 2) bro_tl_st16(t, b, (uint32_t)(uint16_t)v);
            else bro_tl_st32(t, b, (uint32_t)v);
            return *this;
        }
        BRO_MFN const Ref& operator=(const Ref& o) const { return *this = (T)o; }
    };
    BRO_MFN Ref operator[](uint32_t i) const { Ref r; r.t = t; r.b = OFF + i * (uint32_t)sizeof(T); return r; }
};

#ifndef BRO_RING_WORDS
#define BRO_RING_WORDS 6u
#endif
// Layout of a block.  While a meta-block HEADER is read it holds the table reader's scratch; inside a meta-block the
// same bytes hold the decode tables of the current block types (bro_parse.h).  `word` is valid in both.
#define BRO_TL_LENS 0u          // 704 code lengths, 4 bits each (bro_lens_*); the 256-byte IMTF list once they are dead
#define BRO_TL_SYMS 352u        // uint16[4]
#define BRO_TL_CNT 360u         // uint16[16]
#define BRO_TL_LIMIT 392u       // uint16[16]
#define BRO_TL_BASE 424u        // int16[16]
#define BRO_TL_CLC 456u         // uint8[32]
#define BRO_TL_CL 488u          // uint8[18] (+2)
#define BRO_TL_WORD 512u        // uint8[64]
#define BRO_TL_RING 576u        // uint32[BRO_RING_WORDS]: the compressed words behind the bit window (BroBits), always live
#define BRO_TL_BYTES (576u + 4u * BRO_RING_WORDS)
struct BroScratch {
    BroTl t;
    BroTlArray<uint8_t, BRO_TL_LENS> mtf;      // 256-entry move-to-front list (the code lengths are dead by then)
    BroTlArray<uint16_t, BRO_TL_SYMS> syms;    // explicit symbols of a simple code
    BroTlArray<uint16_t, BRO_TL_CNT> cnt;      // per-length running positions
    BroTlArray<uint16_t, BRO_TL_LIMIT> limit;
    BroTlArray<int16_t, BRO_TL_BASE> base;
    BroTlArray<uint8_t, BRO_TL_CLC> clc;       // code-length-code table: symbol | len<<5, indexed by 5 stream bits
    BroTlArray<uint8_t, BRO_TL_CL> cl;         // lengths of the code-length code
    BroTlArray<uint8_t, BRO_TL_WORD> word;     // dictionary word staging (<= 24 + 13 bytes)
    uint16_t *root_lit, *root_cmd, *root_dist; // 256-entry root tables of the fused loops (host simulation only)
    uint8_t *hot_cmap_l, *hot_cmap_d, *hot_modes;   // the general loop's on-chip tables (BRO_HOT_*; host simulation only)
    uint16_t* hot_lit;
};
BRO_FN void bro_scratch_bind(BroScratch& sc, BroTl t) {
    sc.t = t; sc.mtf.t = t; sc.syms.t = t; sc.cnt.t = t; sc.limit.t = t; sc.base.t = t; sc.clc.t = t; sc.cl.t = t; sc.word.t = t;
    sc.root_lit = sc.root_cmd = sc.root_dist = 0;
    sc.hot_cmap_l = sc.hot_cmap_d = sc.hot_modes = 0; sc.hot_lit = 0;
}
// Out-of-line functions get the block's address by value and bind their own view of it (registers), and they work on
// a register copy of the bit window: a struct passed by reference lives in LOCAL memory for its whole life, and with
// hundreds of resident threads per SM every access to it would be an L2 round trip.
#define BRO_SC_PARAM BroTl sc_tl_
#define BRO_SC_BIND BroScratch sc; bro_scratch_bind(sc, sc_tl_)
#define BRO_SC_PASS(sc) (sc).t
#define BRO_CL(sc, i) (sc).cl[i]
#define BRO_CL_DECL
// code lengths (0..15), eight to a word
BRO_FN uint32_t bro_lens_word(const BroScratch& sc, uint32_t w) { return bro_tl_ld32(sc.t, BRO_TL_LENS + 4u * w); }
BRO_FN uint32_t bro_lens_get(const BroScratch& sc, uint32_t i) { return (bro_lens_word(sc, i >> 3) >> (4u * (i & 7u))) & 15u; }
BRO_FN void bro_lens_put(const BroScratch& sc, uint32_t i, uint32_t v) {     // read-modify-write: the few explicit lengths of a simple code
    const uint32_t sh = 4u * (i & 7u);
    bro_tl_st32(sc.t, BRO_TL_LENS + 4u * (i >> 3), (bro_lens_word(sc, i >> 3) & ~(15u << sh)) | (v << sh));
}
// sequential writer: lengths arrive in ascending index order (with gaps that stay zero); a word is stored once
struct BroLensWriter { uint32_t acc, w; };
BRO_FN void bro_lens_begin(const BroScratch& sc, BroLensWriter& lw, uint32_t alphabet) {
    for (uint32_t w = 0; w < (alphabet + 7u) >> 3; w++) bro_tl_st32(sc.t, BRO_TL_LENS + 4u * w, 0u);
    lw.acc = 0; lw.w = 0;
}
BRO_FN void bro_lens_push(const BroScratch& sc, BroLensWriter& lw, uint32_t i, uint32_t v) {
    const uint32_t w = i >> 3;
    if (w != lw.w) { bro_tl_st32(sc.t, BRO_TL_LENS + 4u * lw.w, lw.acc); lw.acc = 0; lw.w = w; }
    lw.acc |= v << (4u * (i & 7u));
}
// a run of `count` lengths v from index i on (repeat code 16: up to the whole alphabet), eight lengths per trip
BRO_FN void bro_lens_push_run(const BroScratch& sc, BroLensWriter& lw, uint32_t i, uint32_t count, uint32_t v) {
    const uint32_t pat = v * 0x11111111u;
    while (count) {
        const uint32_t w = i >> 3, o = i & 7u;
        if (w != lw.w) { bro_tl_st32(sc.t, BRO_TL_LENS + 4u * lw.w, lw.acc); lw.acc = 0; lw.w = w; }
        const uint32_t n = 8u - o < count ? 8u - o : count;
        lw.acc |= (pat & (0xffffffffu >> (4u * (8u - n)))) << (4u * o);
        i += n; count -= n;
    }
}
BRO_FN void bro_lens_end(const BroScratch& sc, BroLensWriter& lw) { bro_tl_st32(sc.t, BRO_TL_LENS + 4u * lw.w, lw.acc); }
#else
// per-warp on-chip scratch (shared memory on the device)
struct BroScratch {
    uint8_t lens[BRO_ALPHA_CMD];   // code lengths of the code being read; reused as the IMTF list
    uint16_t syms[4];              // explicit symbols of a simple code
    uint16_t cnt[16];              // per-length counts, then running positions
    uint16_t limit[16];
    int16_t base[16];
    uint8_t clc[32];               // code-length-code table: symbol | len<<5, indexed by 5 stream bits
    uint8_t word[64];              // dictionary word staging (<= 24 + 13 bytes)
    // on-chip copies of the root tables of the current meta-block's literal / insert&copy / distance code when the
    // meta-block has exactly one of that kind (true for every stream libbrotli produces at quality <= 9)
    uint16_t root_lit[256];
    uint16_t root_cmd[256];
    uint16_t root_dist[256];
    // a meta-block WITH context modelling (several literal / distance codes chosen per symbol, block switches): what its
    // command loop looks up per symbol -- context modes, both context maps, and narrow roots of all its literal codes
    // (of its insert&copy and distance codes in root_cmd / root_dist) -- as far as it fits (bro_stage_hot)
    uint8_t hot_cmap_l[BRO_HOT_CMAP_L];
    uint8_t hot_cmap_d[BRO_HOT_CMAP_D];
    uint8_t hot_modes[BRO_HOT_MODES];
    uint16_t hot_lit[BRO_HOT_LIT_U16];
};
BRO_FN uint32_t bro_lens_get(const BroScratch& sc, uint32_t i) { return sc.lens[i]; }
BRO_FN void bro_lens_put(BroScratch& sc, uint32_t i, uint32_t v) { sc.lens[i] = (uint8_t)v; }
#define BRO_SC_PARAM BroScratch& sc
#define BRO_SC_BIND
#define BRO_SC_PASS(sc) (sc)
#define BRO_CL(sc, i) cl_[i]
#define BRO_CL_DECL uint32_t cl_[18]
#endif

// ------------------------------------------------------------------------------------------------------
// bit stream: src/bitreader/mod.rs:21-304 restated as a 64-bit LSB-first window (w0, w1; the next bit is bit `bp` of w0).
// ------------------------------------------------------------------------------------------------------
#if defined(BRO_SERIAL)
// ONE THREAD PER STREAM (and the host simulation of it).  The words behind the window wait in a small RING in the
// thread's on-chip block (BRO_TL_RING), filled by asynchronous copies global -> shared (cp.async, 4 bytes): a slide takes
// its word from the ring (a shared-memory load) and requests the word BRO_RING_WORDS positions ahead into the slot it
// freed.  No register ever holds data in flight -- in round 1 (one word of lookahead in a register) 17 % of the parse
// kernel's stall samples sat on register moves of words that had not arrived, and with 32 unrelated streams per warp
// some lane slides in nearly every step, so whatever a slide costs is paid all the time: it has to be cheap, uniform
// (predicated, no branch) and never exposed to memory latency.  The sector two ahead is asked of DRAM at the same
// time (prefetch.global.L2).  `avail` counts ALL real bits from `bp` to the end of the stream (streams of 256 MiB
// and more are left to the fused kernel), so that a slide needs no bookkeeping.
struct BroBits {
    const uint8_t* base;    // 4-byte aligned address at or below the first byte of the stream
    const uint8_t* end;     // one past the last byte of the stream
#if defined(BRO_HOSTSIM)
    const uint8_t* lo;      // first byte of the stream (host buffers are neither padded nor aligned)
#endif
    BroTl ring;             // the thread's block (set once per thread, before bro_bits_init)
    uint32_t elen;          // end - base
    uint32_t last;          // offset of the word that holds the stream's last byte
    uint32_t pos;           // offset (from base) of the word that will be requested next; the ring holds the BRO_RING_WORDS words before it
    uint32_t ri;            // ring slot of the word that follows w2
    uint32_t w0, w1;        // the window
    uint32_t w2;            // the word behind it: a slide takes ITS word from the ring one slide before the window needs it, so
                            // that the shared-memory load is never waited for
    uint32_t bp;            // 0..31 after bro_refill
    uint32_t avail;
};

// word at offset o, clamped to the word that holds the stream's last byte: nothing behind that word is ever read.
// What a word holds behind the end of the stream (neighbouring bytes of the same buffer, or a repeat of the last
// word) never matters: `avail` keeps every decision on real bits.
BRO_FN uint32_t bro_word_offset(const BroBits& s, uint32_t o) { return o < s.last ? o : s.last; }
BRO_FN uint32_t bro_load_word(const BroBits& s, uint32_t o) {
#if defined(BRO_HOSTSIM)
    uint32_t w = 0;
    o = bro_word_offset(s, o);
    for (uint32_t i = 0; i < 4u; i++) {
        const uint8_t* p = s.base + o + i;
        if (p >= s.lo && p < s.end) w |= (uint32_t)*p << (8u * i);
    }
    return w;
#else
    return __ldg((const uint32_t*)(s.base + bro_word_offset(s, o)));
#endif
}

// position the window at byte address `a` (start of stream, or after a stored / metadata block)
BRO_FN void bro_bits_seek(BroBits& s, const uint8_t* a) {
    const uint32_t off = (uint32_t)(a - s.base), wo = off & ~3u;
    s.ri = 0;
    s.pos = wo + 12u + 4u * BRO_RING_WORDS;
    s.bp = 8u * (off & 3u);
    s.avail = a < s.end ? 8u * (uint32_t)(s.end - a) : 0u;
    if (s.elen == 0u) { s.w0 = s.w1 = s.w2 = 0; return; }     // an empty stream: nothing to read, and nothing is ever consumed
#if defined(BRO_WARPSIM)
    ws_ring_wait_group(0u);
#elif !defined(BRO_HOSTSIM)
    asm volatile("cp.async.wait_all;");                         // no request of the old position may land in a slot later
#endif
    s.w0 = bro_load_word(s, wo);
    s.w1 = bro_load_word(s, wo + 4u);
    s.w2 = bro_load_word(s, wo + 8u);
    for (uint32_t j = 0; j < BRO_RING_WORDS; j++) {
#if defined(BRO_HOSTSIM)
        bro_tl_st32(s.ring, BRO_TL_RING + 4u * j, bro_load_word(s, wo + 12u + 4u * j));
#elif defined(BRO_WARPSIM)
        ws_ring_issue(s.ring.base + ((BRO_TL_RING + 4u * j) << 5), s.base + bro_word_offset(s, wo + 12u + 4u * j));
#else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n\tcp.async.commit_group;"
                     :: "r"(s.ring.base + ((BRO_TL_RING + 4u * j) << 5)), "l"(s.base + bro_word_offset(s, wo + 12u + 4u * j)));
#endif
    }
}

BRO_FN void bro_bits_init(BroBits& s, const uint8_t* start, const uint8_t* end) {
#if defined(BRO_HOSTSIM)
    s.lo = start;
#endif
    s.base = (const uint8_t*)((uintptr_t)start & ~(uintptr_t)3);
    s.end = end;
    s.elen = (uint32_t)(end - s.base);
    s.last = (s.elen - 1u) & ~3u;
    bro_bits_seek(s, start);
}

// Protocol: bro_refill() (slides the window so that bp < 32), then bro_peek() / bro_avail(), then bro_consume() of at
// most 32 bits.  A lane slides at most once per call, and every call is one cp.async group (of the lanes that slid), so
// the word a lane takes from the ring was requested at least BRO_RING_WORDS groups ago, however the hardware counts
// groups (per thread or per warp): waiting for all but the BRO_RING_WORDS - 1 youngest is enough.  One predicated
// block, no branch: wait, take the word, request the word BRO_RING_WORDS ahead into the freed slot.
BRO_FN void bro_refill(BroBits& s) {
#if defined(BRO_HOSTSIM)
    if (s.bp >= 32u) {
        s.w0 = s.w1;
        s.w1 = s.w2;
        s.w2 = bro_tl_ld32(s.ring, BRO_TL_RING + 4u * s.ri);
        bro_tl_st32(s.ring, BRO_TL_RING + 4u * s.ri, bro_load_word(s, s.pos));
        s.pos += 4u;
        s.ri = s.ri + 1u == BRO_RING_WORDS ? 0u : s.ri + 1u;
        s.bp -= 32u;
    }
#elif defined(BRO_WARPSIM)
    // the block below, statement by statement; the asynchronous copies land as late as the wait allows (bro_warpsim_parse.cpp)
    if (s.bp >= 32u) {
        const uint32_t slot = s.ring.base + (BRO_TL_RING << 5) + s.ri * 128u;
        ws_ring_wait_group(BRO_RING_WORDS - 1u);
        s.w0 = s.w1;
        s.w1 = s.w2;
        s.w2 = *(const uint32_t*)ws_smem_ptr(slot);
        ws_ring_issue(slot, s.base + (s.pos < s.last ? s.pos : s.last));
        s.pos += 4u;
        s.ri += 1u;
        s.bp -= 32u;
        if (s.ri == BRO_RING_WORDS) s.ri = 0u;
    }
#else
    asm volatile("{\n\t"
                 ".reg .pred p, w;\n\t"
                 ".reg .u32 o, slot;\n\t"
                 ".reg .u64 a;\n\t"
                 "setp.ge.u32 p, %3, 32;\n\t"
                 "mad.lo.u32 slot, %5, 128, %6;\n\t"          // the slot of word ri: interleaved, 128 bytes apart
                 "min.u32 o, %4, %8;\n\t"
                 "cvt.u64.u32 a, o;\n\t"
                 "add.u64 a, a, %7;\n\t"
                 "@p cp.async.wait_group %9;\n\t"
                 "@p mov.u32 %0, %1;\n\t"
                 "@p mov.u32 %1, %2;\n\t"
                 "@p ld.shared.u32 %2, [slot];\n\t"
                 "@p cp.async.ca.shared.global [slot], [a], 4;\n\t"
                 "@p cp.async.commit_group;\n\t"
                 "@p add.u32 %4, %4, 4;\n\t"
                 "@p add.u32 %5, %5, 1;\n\t"
                 "@p sub.u32 %3, %3, 32;\n\t"
                 "setp.eq.and.u32 w, %5, %10, p;\n\t"
                 "@w mov.u32 %5, 0;\n\t"
                 "}"
                 : "+r"(s.w0), "+r"(s.w1), "+r"(s.w2), "+r"(s.bp), "+r"(s.pos), "+r"(s.ri)
                 : "r"(s.ring.base + (BRO_TL_RING << 5)), "l"(s.base), "r"(s.last), "n"(BRO_RING_WORDS - 1u), "n"(BRO_RING_WORDS));
#endif
}
// byte offset (from base) of window word w0
BRO_FN uint32_t bro_bits_w0(const BroBits& s) { return s.pos - 4u * BRO_RING_WORDS - 12u; }
#else
// ONE WARP (or lane group) PER STREAM: the group loads the compressed bytes 32 words at a time (lane i holds word i,
// the next chunk is already in flight) and feeds the window by shuffle.
struct BroBits {
    const uint8_t* chunk;   // aligned address of the chunk held in `cur`
    const uint8_t* lo;      // first loadable word address (stream start rounded down to 4)
    const uint8_t* end;     // one past the last byte of the stream
    uint32_t w0, w1;        // bit window: two consecutive little-endian words of the stream; next bit = bit `bp` of w0
    uint32_t bp;            // 0..31
    uint32_t avail;         // real stream bits in the window from `bp` on (the rest of w0/w1 is padding past the end)
    uint32_t rem;           // real stream BYTES not yet loaded into the window (a stream is < 4 GiB)
    uint32_t wi;            // next word of the chunk to hand out
    uint32_t cur, nxt;      // this lane's word of the current / next chunk
};

BRO_FN uint32_t bro_load_word(const BroBits& s, const uint8_t* a) {
    // words entirely outside [lo, end) read as zero; words straddling the ends expose neighbouring bytes of the
    // same allocation, which the bit accounting below never lets a decision depend on
    if (a < s.lo || a >= s.end) return 0u;
    return __ldg((const uint32_t*)a);
}

BRO_FN uint32_t bro_next_word(BroBits& s) {
    uint32_t w = bro_shfl(s.cur, s.wi);
    if (++s.wi == BRO_W) {
        s.cur = s.nxt;
        s.chunk += 4u * BRO_W;
        s.nxt = bro_load_word(s, s.chunk + 4u * BRO_W + 4u * bro_lane());
        s.wi = 0;
    }
    return w;
}

// position the window at byte address `a` (start of stream, or after a stored / metadata block)
BRO_FN void bro_bits_seek(BroBits& s, const uint8_t* a) {
    uintptr_t ai = (uintptr_t)a;
    s.chunk = (const uint8_t*)(ai & ~(uintptr_t)(4u * BRO_W - 1u));
    s.wi = (uint32_t)(ai & (4u * BRO_W - 1u)) >> 2;
    s.cur = bro_load_word(s, s.chunk + 4u * bro_lane());
    s.nxt = bro_load_word(s, s.chunk + 4u * BRO_W + 4u * bro_lane());
    uint32_t left = a < s.end ? (uint32_t)(s.end - a) : 0u;      // real bytes from `a` on
    uint32_t sh = (uint32_t)(ai & 3u);
    s.w0 = bro_next_word(s);
    s.w1 = bro_next_word(s);
    s.bp = 8u * sh;
    uint32_t in_window = 8u - sh;                                 // bytes of [a, ...) the two words cover
    if (in_window > left) in_window = left;
    s.avail = 8u * in_window;
    s.rem = left - in_window;
}

BRO_FN void bro_bits_init(BroBits& s, const uint8_t* start, const uint8_t* end) {
    s.lo = (const uint8_t*)((uintptr_t)start & ~(uintptr_t)3);
    s.end = end;
    bro_bits_seek(s, start);
}

// Protocol: bro_refill() (slides the window so that bp < 32), then bro_peek() / bro_avail(), then bro_consume() of at
// most 32 bits.  Keeping the slide in ONE place per read keeps the hot loops small (the I-cache is a first-order
// limit for this kernel).
BRO_FN void bro_refill(BroBits& s) {
    if (s.bp >= 32u) {
        s.bp -= 32u;
        s.w0 = s.w1;
        s.w1 = bro_next_word(s);
        uint32_t got = s.rem < 4u ? s.rem : 4u;
        s.avail += 8u * got;
        s.rem -= got;
    }
}
#endif
// The next 32 stream bits, first bit in bit 0 (after bro_refill the window holds more than 32 bits past `bp`).
BRO_FN uint32_t bro_peek(const BroBits& s) { return bro_funnel_r(s.w0, s.w1, s.bp); }
// The same when up to 17 bits have been consumed since bro_refill (bp < 49): at least 15 valid bits.
BRO_FN uint32_t bro_peek_wide(const BroBits& s) {
    const bool hi = s.bp >= 32u;
    return bro_funnel_r(hi ? s.w1 : s.w0, hi ? 0u : s.w1, s.bp & 31u);
}
BRO_FN uint32_t bro_avail(const BroBits& s) { return s.avail; }
BRO_FN void bro_consume(BroBits& s, uint32_t n) { s.bp += n; s.avail -= n; }   // n <= 32, n <= avail

// n <= 25 bits, least significant first (src/bitreader/mod.rs:140-158).  Returns false at end of input, which
// every caller in the reference maps to UnexpectedEOF.
BRO_FN bool bro_read_bits(BroBits& s, uint32_t n, uint32_t& v) {
    bro_refill(s);
    if (n > s.avail) return false;
    v = bro_peek(s) & ((1u << n) - 1u);
    bro_consume(s, n);
    return true;
}

// src/bitreader/mod.rs:257-267: the bits up to the next byte boundary (0 if already aligned)
BRO_FN bool bro_read_byte_tail(BroBits& s, uint32_t& v) {
    bro_refill(s);
    uint32_t n = s.avail & 7u;   // real bits left in the stream are a whole number of bytes plus the tail
    v = bro_peek(s) & ((1u << n) - 1u);
    bro_consume(s, n);
    return true;
}

// byte address of the next unread bit (valid when byte aligned)
BRO_FN const uint8_t* bro_bits_addr(const BroBits& s) {
#if defined(BRO_SERIAL)
    return s.base + (bro_bits_w0(s) + (s.bp >> 3));
#else
    return s.chunk + 4u * s.wi - 8 + (s.bp >> 3);
#endif
}

// bits consumed since byte address `start` (any state of the window)
BRO_FN uint64_t bro_bits_position(const BroBits& s, const uint8_t* start) {
#if defined(BRO_SERIAL)
    const uint8_t* w0 = s.base + (int32_t)bro_bits_w0(s);       // (below base right after a seek to the stream's first bytes)
#else
    const uint8_t* w0 = s.chunk + 4u * s.wi - 8;
#endif
    return (uint64_t)((int64_t)(w0 - start) * 8 + (int64_t)s.bp);
}

// ------------------------------------------------------------------------------------------------------
// prefix code tables
// ------------------------------------------------------------------------------------------------------
#define BRO_SYM_OK 0
#define BRO_SYM_EOF 1
#define BRO_SYM_HOLE 2

// Decode one symbol (src/huffman/tree/mod.rs:63-92 restated).  The reference walks one bit at a time and
// stops at the first assigned node; it reports Ok(None) only after max_depth+1 bits if no node was hit (SURVEY
// Q5), and a failed bit read before that is an EOF.  A table hit whose code is longer than the remaining input
// is therefore EOF, and a hole is EOF unless max_depth+1 real bits remain.
// Codes longer than 8 bits, single-symbol tables and holes: out of line, and it does not touch the window.
// Returns symbol | length << 16 | BRO_SYM_* << 24.
BRO_COLD uint32_t bro_sym_slow(const uint16_t* T, uint32_t peek, uint32_t e, uint32_t avail, uint32_t root_bits = BRO_ROOT_BITS) {
    if (T[BRO_T_SINGLE]) return (uint32_t)T[BRO_T_SINGLE_SYM];
    if (e == 1u) {
        uint32_t x = bro_brev(peek) >> 17;   // next 15 bits, first bit read most significant
        uint32_t L = root_bits + 1u;        // no code of at most root_bits bits starts here
        while (L <= 15u && x >= T[BRO_T_LIMIT + L]) L++;
        if (L <= 15u) {
            if (L > avail) return (uint32_t)BRO_SYM_EOF << 24;
            uint32_t sym = T[BRO_T_SORTED + (int)(int16_t)T[BRO_T_BASE + L] + (int)(x >> (15u - L))];
            return sym | (L << 16);
        }
    }
    return (uint32_t)((avail >= (uint32_t)T[BRO_T_MAXDEPTH] + 1u) ? BRO_SYM_HOLE : BRO_SYM_EOF) << 24;
}

// `root` is the 256-entry root table of T: T itself (HBM, L1-cached) or its on-chip copy in the scratch.
BRO_FN int bro_decode_sym2(BroBits& s, const uint16_t* root, const uint16_t* T, uint32_t& sym) {
    bro_refill(s);
    uint32_t peek = bro_peek(s);
    uint32_t e = root[peek & (BRO_ROOT_SIZE - 1u)];
    uint32_t len = e >> 10;
    if (len != 0u) {
        if (len > bro_avail(s)) return BRO_SYM_EOF;
        bro_consume(s, len);
        sym = e & 0x3ffu;
        return BRO_SYM_OK;
    }
    uint32_t r = bro_sym_slow(T, peek, e, bro_avail(s));
    bro_consume(s, (r >> 16) & 0xffu);
    sym = r & 0xffffu;
    return (int)(r >> 24);
}

BRO_FN int bro_decode_sym(BroBits& s, const uint16_t* T, uint32_t& sym) { return bro_decode_sym2(s, T, T, sym); }

// Build a tree record from n (length, symbol) pairs in sc.lens[] (and sc.syms[] when `explicit_syms`), in
// the order the reference inserts them: src/huffman/mod.rs:19-43 assigns canonical codes per length in array
// order; Tree::insert (src/huffman/tree/mod.rs:50-61) counts every insert, and a tree with exactly one
// insert decodes with zero bits.  The 32 lanes cooperate: counting sort by length, then every lane resolves
// 8 of the 256 root entries by searching the canonical limits.
#if defined(BRO_SERIAL)
// One thread builds the record (thread-per-stream parse kernel, host simulation).  Everything the thread READS here is
// in its own scratch (local memory on the device); the table itself is only written: the root is filled by
// replication as every symbol is placed (its canonical code is known at that moment), so the build never waits for a
// load from the table arena in HBM.
// want_root = false: the caller decodes this table canonically (limits, bases and sorted[] only): the 256-entry root
// is neither cleared nor filled.
BRO_COLD void bro_build_tree(uint16_t* T, BRO_SC_PARAM, uint32_t n, bool explicit_syms, bool want_root = true) {
    BRO_SC_BIND;
    // pass 1: per-length counts (running positions later) in the thread's on-chip block; unused symbols -- most of the
    // 704 insert&copy symbols of a typical code -- are skipped eight at a time
    for (uint32_t L = 0; L < 16u; L++) sc.cnt[L] = 0;
    const uint32_t nw = (n + 7u) >> 3;
    for (uint32_t w = 0; w < nw; w++) {
        uint32_t word = bro_lens_word(sc, w);
        if (8u * w + 8u > n) word &= (1u << (4u * (n - 8u * w))) - 1u;         // lengths behind the alphabet are not part of the code
        while (word) {
            const uint32_t L = word & 15u;
            word >>= 4;
            if (L) sc.cnt[L] = (uint16_t)(sc.cnt[L] + 1u);
        }
    }
    uint32_t code = 0, off = 0, maxdepth = 0, nonzero = 0;
    T[BRO_T_LIMIT] = 0; T[BRO_T_BASE] = 0;
    for (uint32_t L = 1; L <= 15u; L++) {
        const uint32_t c = sc.cnt[L];
        sc.cnt[L] = (uint16_t)off;                                            // next free position of this length in sorted[]
        const int base = (int)off - (int)code;                                // canonical index of the first code of the length minus its code value
        sc.base[L] = (int16_t)base;
        T[BRO_T_BASE + L] = (uint16_t)(int16_t)base;
        code += c;
        T[BRO_T_LIMIT + L] = (uint16_t)(code << (15u - L));
        code <<= 1;
        off += c;
        nonzero += c;
        if (c) maxdepth = L;
    }
    // all-zero lengths happen only for a simple code with NSYM = 1 (src/huffman/mod.rs:36); a complex code always has
    // >= 2 non-zero lengths
    T[BRO_T_SINGLE] = (nonzero <= 1u) ? 1 : 0;
    T[BRO_T_MAXDEPTH] = (uint16_t)maxdepth;
    T[BRO_T_MAXDEPTH + 1u] = 0;
#if defined(BRO_HOSTSIM)
    if (want_root) for (uint32_t r = 0; r < BRO_ROOT_SIZE; r++) T[r] = 0;
#else
    if (want_root) for (uint32_t r = 0; r < BRO_ROOT_SIZE / 8u; r++) ((uint4*)T)[r] = make_uint4(0u, 0u, 0u, 0u);     // records are 16-byte aligned
#endif
    // pass 2: place the symbols in canonical order (array order inside a length); the table itself is only written --
    // the root is filled by replication as every symbol is placed (its canonical code is known at that moment)
    uint32_t single_sym = explicit_syms ? (uint32_t)sc.syms[0] : 0u;          // nonzero == 0: the first pair
    for (uint32_t w = 0; w < nw; w++) {
        uint32_t word = bro_lens_word(sc, w);
        if (8u * w + 8u > n) word &= (1u << (4u * (n - 8u * w))) - 1u;
        for (uint32_t i = 8u * w; word; i++, word >>= 4) {
            const uint32_t L = word & 15u;
            if (L == 0u) continue;
            const uint32_t symv = explicit_syms ? (uint32_t)sc.syms[i] : i;
            const uint32_t idx = sc.cnt[L];
            sc.cnt[L] = (uint16_t)(idx + 1u);
            T[BRO_T_SORTED + idx] = (uint16_t)symv;
            if (nonzero == 1u) single_sym = symv;
            if (nonzero >= 2u && want_root) {
                const uint32_t c = (uint32_t)((int)idx - (int)sc.base[L]);    // the canonical code of this symbol
                if (L <= BRO_ROOT_BITS) {
                    const uint32_t e = symv | (L << 10);
                    for (uint32_t r = bro_brev(c) >> (32u - L); r < BRO_ROOT_SIZE; r += 1u << L) T[r] = (uint16_t)e;
                } else T[bro_brev(c >> (L - BRO_ROOT_BITS)) >> (32u - BRO_ROOT_BITS)] = 1;
            }
        }
    }
    T[BRO_T_SINGLE_SYM] = (uint16_t)single_sym;
}
#else
BRO_COLD void bro_build_tree(uint16_t* T, BroScratch& sc, uint32_t n, bool explicit_syms, bool = true) {
    const unsigned lane = bro_lane();
    for (unsigned i = lane; i < 16u; i += BRO_W) sc.cnt[i] = 0;
    bro_syncwarp();
    // pass 1: histogram of lengths (leaders of equal-length groups add the group size)
    for (uint32_t i0 = 0; i0 < n; i0 += BRO_W) {
        uint32_t i = i0 + lane;
        uint32_t L = i < n ? sc.lens[i] : 0xffu;
        uint32_t m = bro_match_any(L);
        if (i < n && (m & bro_lanemask_lt()) == 0u) sc.cnt[L] = (uint16_t)(sc.cnt[L] + bro_popc(m));
        bro_syncwarp();
    }
    // canonical first codes, left-justified limits, sorted[] offsets (uniform, 15 steps)
    uint32_t code = 0, off = 0, maxdepth = 0, nonzero = 0;
    uint32_t first[16], offs[16];
    for (uint32_t L = 1; L <= 15u; L++) {
        uint32_t c = sc.cnt[L];
        first[L] = code;
        offs[L] = off;
        code = (code + c);
        uint32_t lim = code << (15u - L);
        code <<= 1;
        off += c;
        nonzero += c;
        if (c) maxdepth = L;
        if (lane == 0) {
            sc.limit[L] = (uint16_t)lim;
            sc.base[L] = (int16_t)((int)offs[L] - (int)first[L]);
        }
    }
    bro_syncwarp();
    if (lane == 0) {
        for (uint32_t L = 1; L <= 15u; L++) { sc.cnt[L] = (uint16_t)offs[L]; T[BRO_T_LIMIT + L] = sc.limit[L]; T[BRO_T_BASE + L] = (uint16_t)sc.base[L]; }
        T[BRO_T_LIMIT] = 0; T[BRO_T_BASE] = 0;
        // all-zero lengths happen only for a simple code with NSYM = 1 (max_length == 0: every symbol is inserted
        // with the empty code, src/huffman/mod.rs:36); a complex code always has >= 2 non-zero lengths
        bool single = (nonzero == 1u) || (nonzero == 0u);
        T[BRO_T_SINGLE] = single ? 1 : 0;
        T[BRO_T_MAXDEPTH] = (uint16_t)maxdepth;
        T[BRO_T_MAXDEPTH + 1u] = 0;
    }
    bro_syncwarp();
    // pass 2: stable placement into sorted[] (rank inside the 32-symbol group + running position per length)
    for (uint32_t i0 = 0; i0 < n; i0 += BRO_W) {
        uint32_t i = i0 + lane;
        uint32_t L = i < n ? sc.lens[i] : 0xffu;
        uint32_t m = bro_match_any(L);
        bool live = i < n && L != 0u;
        uint32_t symv = explicit_syms ? (i < n ? sc.syms[i] : 0u) : i;
        if (live) T[BRO_T_SORTED + sc.cnt[L] + bro_popc(m & bro_lanemask_lt())] = (uint16_t)symv;
        if (nonzero == 0u && i == 0u) T[BRO_T_SINGLE_SYM] = (uint16_t)symv;
        bro_syncwarp();
        if (live && (m & bro_lanemask_lt()) == 0u) sc.cnt[L] = (uint16_t)(sc.cnt[L] + bro_popc(m));
        bro_syncwarp();
    }
    if (nonzero == 1u && lane == 0) T[BRO_T_SINGLE_SYM] = T[BRO_T_SORTED];
    // root table: entry r is reached when the next 8 stream bits, read LSB first, equal r
    for (uint32_t r = lane; r < 256u; r += BRO_W) {
        uint32_t v = bro_brev(r) >> 24;          // the same 8 bits, first bit read most significant
        uint32_t x = v << 7;
        uint32_t e = 0;
        if (nonzero >= 2u) {
            for (uint32_t L = 1; L <= 15u; L++) {
                if (x < sc.limit[L]) {
                    if (L <= 8u) e = (uint32_t)T[BRO_T_SORTED + (int)sc.base[L] + (int)(v >> (8u - L))] | (L << 10);
                    else e = 1u;
                    break;
                }
            }
        }
        T[r] = (uint16_t)e;
    }
    bro_syncwarp();
}

#endif

// ------------------------------------------------------------------------------------------------------
// decoder state
// ------------------------------------------------------------------------------------------------------
struct BroBlockCat {          // src/lib.rs:152-160, one per category (literals, insert&copy, distances)
    uint32_t nbl;
    uint32_t btype, btype_prev;
    uint32_t blen;            // valid when nbl >= 2 (the reference's Option<BLen> is Some exactly then)
    uint32_t t_type, t_count; // arena offsets of the block-type and block-count code tables
};

struct BroDec {
    BroBits in;
    uint8_t* out;             // output slot
    uint32_t cap;             // slot capacity
    uint32_t pos;             // bytes produced so far (= count_output, src/lib.rs:385)
    uint32_t window;          // (1 << WBITS) - 16, src/lib.rs:1562
    uint32_t p1, p2;          // literal_buf, src/lib.rs:389
    uint32_t d0, d1, d2, d3;  // distance_buf, src/lib.rs:393; d0 is the last distance
    uint16_t* arena;          // table arena of this warp / thread (bump-allocated per meta-block)
    uint32_t arena_cap;       // its capacity in uint16 units
    uint32_t arena_base;      // first free uint16 (thread mode keeps its scratch below it)
#if defined(BRO_SERIAL)
    BroScratch scv;           // view of this thread's on-chip block (a few addresses, by value: registers)
#define BRO_DSC(d) ((d).scv)
#else
    BroScratch* sc;           // this warp's scratch in shared memory
    uint16_t* root10;         // latency build only (else 0): 1,024 entries of shared memory for the 10-bit literal root of the
                              // general loop's lane-parallel rounds (bro_lit_row).  The throughput build runs 24 warps per SM
                              // and needs what is left of the SM's 256 KB as L1 for its tables: 2 KB more per warp cost it 5 %
#define BRO_DSC(d) (*(d).sc)
#endif
    const uint8_t* dict;      // 122,784-byte static dictionary image
    int quirk_spec;
#if defined(BRO_PARSE)
    BroRec* rec;              // copy records of this stream (phase one of the two-phase path writes, phase two executes)
    uint32_t nrec, rec_cap;
    const uint8_t* in_base;   // first byte of the compressed stream (stored-block records hold offsets from it)
    const uint32_t* ic;       // insert / copy length codes -> base | extra bits << 16 (bro_ic_lookup; shared memory)
    uint32_t out_mis;         // (address of out) & 15: pieces are cut at 16-byte boundaries of the destination ADDRESS
    uint32_t sizing;          // 1: only measure the stream (bro_batch_sizes): nothing is written, the slot is unbounded
    uint32_t imm;             // 1: this stream's copies are executed by its own thread as they are decoded (no records): what a
                              // meta-block WITH literal context modelling needs -- the two bytes in front of every literal
#endif
};

#if defined(BRO_PARSE)
BRO_FN bool bro_rec_push1(BroDec& d, uint32_t dst, uint32_t len, uint32_t kind, uint32_t a) {
    if (d.nrec >= d.rec_cap) return false;
    BroRec r;
    r.dst = dst; r.len_kind = len | (kind << BRO_REC_KIND_SHIFT); r.a = a; r.b = 0;
#if defined(BRO_THREAD_MODE)
    // written once, read once by the next kernel: streaming store, so that records do not push the prefix tables
    // (looked up for every symbol) out of L2
    __stcs((uint4*)(d.rec + d.nrec), make_uint4(r.dst, r.len_kind, r.a, r.b));
    d.nrec++;
#else
    d.rec[d.nrec++] = r;
#endif
    return true;
}

// A copy becomes one record per PIECE of at most BRO_REC_PIECE_VECS 16-byte vectors on 16-byte aligned destinations
// (plus the ragged bytes before the first and after the last vector), so that one step of the copy kernel's warp moves
// one piece whatever the copy's length.  Pieces keep the copy's distance: a piece may read what an earlier piece of
// the same copy wrote, which the copy kernel's grouping handles like any other dependency.  Only a copy whose distance
// is shorter than a piece (a periodic fill) stays whole.
BRO_FN bool bro_rec_push(BroDec& d, uint32_t dst, uint32_t len, uint32_t kind, uint32_t a) {
    if (d.sizing) return true;
    if (kind == BRO_REC_LZ && a < len && a < 16u * BRO_REC_PIECE_VECS + 32u) return bro_rec_push1(d, dst, len, kind, a);
    // first piece: up to the 16-byte boundary plus BRO_REC_PIECE_VECS vectors; then whole pieces; the last few bytes
    // (< 16) ride along as the tail of the piece before them
    uint32_t piece = ((16u - ((d.out_mis + dst) & 15u)) & 15u) + 16u * BRO_REC_PIECE_VECS;
    bool ok = true;
    while (ok && len != 0u) {
        if (len < piece + 16u) piece = len;
        ok = bro_rec_push1(d, dst, piece, kind, a);
        dst += piece; len -= piece;
        if (kind == BRO_REC_STORED) a += piece;
        piece = 16u * BRO_REC_PIECE_VECS;
    }
    return ok;
}
#endif

// The fixed code of NBLTYPES / NTREES (src/lib.rs:126-132, 501-525): 0 -> 1, else 1 + (1 << n) + n extra bits
// with n read from 3 bits.  Any failed bit read is UnexpectedEOF.
BRO_FN int bro_read_nbltypes(BroBits& in, uint32_t& v) {
    uint32_t b, n, extra;
    if (!bro_read_bits(in, 1, b)) return BRO_ST_UnexpectedEOF;
    if (!b) { v = 1; return 0; }
    if (!bro_read_bits(in, 3, n)) return BRO_ST_UnexpectedEOF;
    if (!bro_read_bits(in, n, extra)) return BRO_ST_UnexpectedEOF;
    v = 1u + (1u << n) + extra;
    return 0;
}

// The readers below only fill the scratch's code lengths (and sc.syms[]); the table is built by their caller afterwards.
// In the thread-per-stream kernel this matters: lanes leave the loops below at different trips, and they are converged
// again only once the out-of-line function has returned -- the build then runs once for the whole warp.  For the same
// reason the loops have a single exit (errors break and are returned after the loop).

// src/lib.rs:597-665.  -> number of (length, symbol) pairs in the scratch's lengths / sc.syms[]
BRO_FN int bro_read_simple_code(BroBits& in, BroScratch& sc, uint32_t alphabet, uint32_t& n_out) {
    uint32_t bit_width = 0;
    for (uint32_t a = alphabet - 1u; a; a >>= 1) bit_width++;   // 16 - leading_zeros(alphabet-1 as u16), src/lib.rs:598
    uint32_t nsym, s[4] = {0, 0, 0, 0};
    if (!bro_read_bits(in, 2, nsym)) return BRO_ST_UnexpectedEOF;
    nsym += 1;
    int st = 0;
#pragma unroll
    for (uint32_t i = 0; i < 4u; i++) {
        if (i < nsym && !st) {
            if (!bro_read_bits(in, bit_width, s[i])) st = BRO_ST_UnexpectedEOF;
            else if (s[i] >= alphabet) st = BRO_ST_InvalidSymbol;
        }
    }
    if (st) return st;
#pragma unroll
    for (uint32_t i = 0; i < 3u; i++)
#pragma unroll
        for (uint32_t j = i + 1; j < 4u; j++)
            if (j < nsym && s[i] == s[j]) st = BRO_ST_InvalidSymbol;
    if (st) return st;
    uint32_t L[4] = {0, 0, 0, 0};
#define BRO_SWAP(a, b) do { if (s[a] > s[b]) { uint32_t t_ = s[a]; s[a] = s[b]; s[b] = t_; } } while (0)
    if (nsym == 2) { BRO_SWAP(0, 1); L[0] = L[1] = 1; }
    else if (nsym == 3) { BRO_SWAP(1, 2); L[0] = 1; L[1] = L[2] = 2; }
    else if (nsym == 4) {
        uint32_t tree_select;
        if (!bro_read_bits(in, 1, tree_select)) return BRO_ST_UnexpectedEOF;
        if (!tree_select) {
            BRO_SWAP(0, 1); BRO_SWAP(2, 3); BRO_SWAP(0, 2); BRO_SWAP(1, 3); BRO_SWAP(1, 2);
            L[0] = L[1] = L[2] = L[3] = 2;
        } else { BRO_SWAP(2, 3); L[0] = 1; L[1] = 2; L[2] = L[3] = 3; }
    }
#undef BRO_SWAP
    bro_syncwarp();
    if (bro_lane() == 0) {
#if defined(BRO_SERIAL)
        bro_tl_st32(sc.t, BRO_TL_LENS, L[0] | (L[1] << 4) | (L[2] << 8) | (L[3] << 12));
#endif
#pragma unroll
        for (uint32_t i = 0; i < 4u; i++) if (i < nsym) {
#if !defined(BRO_SERIAL)
            sc.lens[i] = (uint8_t)L[i];
#endif
            sc.syms[i] = (uint16_t)s[i];
        }
    }
    bro_syncwarp();
    n_out = nsym;
    return 0;
}

// src/lib.rs:667-875.  Fills the scratch's code lengths [0, alphabet).
BRO_FN int bro_read_complex_code(BroBits& in, BroScratch& sc, uint32_t hskip, uint32_t alphabet) {
    const unsigned lane = bro_lane();
    // code lengths of the code-length code, transmitted in the order 1,2,3,4,0,5,17,6,16,7,8,...,15 with the fixed
    // code 00->0 01->3 10->4 110->2 1110->1 1111->5 (src/lib.rs:120-125, 669-704)
    BRO_CL_DECL;
    for (uint32_t i = 0; i < 18u; i++) BRO_CL(sc, i) = 0;
    uint32_t sum = 0, nonzero = 0;
    int st = 0;
    for (uint32_t i = hskip; i < 18u; i++) {
        uint32_t b, v;
        if (!bro_read_bits(in, 2, b)) { st = BRO_ST_UnexpectedEOF; break; }
        if (b == 0u) v = 0;
        else if (b == 2u) v = 3;          // read order 0,1
        else if (b == 1u) v = 4;          // read order 1,0
        else {
            if (!bro_read_bits(in, 1, b)) { st = BRO_ST_UnexpectedEOF; break; }
            if (!b) v = 2;
            else {
                if (!bro_read_bits(in, 1, b)) { st = BRO_ST_UnexpectedEOF; break; }
                v = b ? 5 : 1;
            }
        }
        // transmission slot i -> symbol: 1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, ..., 15 (one nibble / one bit per slot)
        const uint32_t slot_sym = i < 16u ? (uint32_t)((0xdcba987061504321ull >> (4u * i)) & 15u) + ((0x0140u >> i) & 1u) * 16u : i - 2u;
        BRO_CL(sc, slot_sym) = v;
        if (v > 0u) {
            sum += 32u >> v;
            nonzero += 1;
            if (sum == 32u) break;
            if (sum > 32u) { st = BRO_ST_CodeLengthsChecksum; break; }
        }
    }
    if (st) return st;
    if (nonzero == 0u) return BRO_ST_NoCodeLength;
    if (nonzero >= 2u && sum < 32u) return BRO_ST_CodeLengthsChecksum;

    // 32-entry table for the code-length code (max length 5), canonical codes by (length, symbol)
    uint32_t clc_single = 0xffffffffu;
    {
        uint32_t code = 0;
        bro_syncwarp();
        for (uint32_t L = 1; L <= 5u; L++) {
            for (uint32_t sy = 0; sy < 18u; sy++) {
                if (BRO_CL(sc, sy) == L) {
                    uint32_t rev = bro_brev(code) >> (32u - L);
#if defined(BRO_SERIAL)
                    for (uint32_t r = rev; r < 32u; r += 1u << L) sc.clc[r] = (uint8_t)(sy | (L << 5));
#else
                    for (uint32_t r = lane; r < 32u; r += BRO_W)
                        if ((r & ((1u << L) - 1u)) == rev) sc.clc[r] = (uint8_t)(sy | (L << 5));
#endif
                    code++;
                }
            }
            code <<= 1;
        }
        if (nonzero == 1u) for (uint32_t sy = 0; sy < 18u; sy++) if (BRO_CL(sc, sy)) clc_single = sy;
        bro_syncwarp();
    }

    // the symbol code lengths (src/lib.rs:730-864)
#if defined(BRO_SERIAL)
    BroLensWriter lw;
    bro_lens_begin(sc, lw, alphabet);
#else
    for (uint32_t i = lane; i < alphabet; i += BRO_W) sc.lens[i] = 0;
    bro_syncwarp();
#endif
    uint32_t total = 0, last_symbol = 0xffu, last_repeat = 0, have_repeat = 0, last_nz = 8, i = 0, nz = 0;
    while (i < alphabet) {
        uint32_t c;
        if (clc_single != 0xffffffffu) c = clc_single;
        else {
            bro_refill(in);
            uint32_t e = sc.clc[bro_peek(in) & 31u];
            uint32_t len = e >> 5;
            if (len > bro_avail(in)) { st = BRO_ST_UnexpectedEOF; break; }
            bro_consume(in, len);
            c = e & 31u;
        }
        if (c <= 15u) {
#if defined(BRO_SERIAL)
            bro_lens_push(sc, lw, i, c);
#else
            if (lane == 0) sc.lens[i] = (uint8_t)c;
#endif
            i += 1;
            last_symbol = c;
            have_repeat = 0;
            if (c > 0u) {
                last_nz = c;
                nz += 1;
                total += 32768u >> c;
                if (total == 32768u) break;
                if (total > 32768u) { st = BRO_ST_CodeLengthsChecksum; break; }
            }
        } else if (c == 16u) {
            uint32_t extra, count, newrep;
            if (!bro_read_bits(in, 2, extra)) { st = BRO_ST_UnexpectedEOF; break; }
            if (last_symbol == 16u && have_repeat) {
                newrep = 4u * (last_repeat - 2u) + extra + 3u;
                if (i + newrep - last_repeat > alphabet) { st = BRO_ST_ParseErrorComplexPrefixCodeLengths; break; }
                count = newrep - last_repeat;
            } else {
                newrep = 3u + extra;
                if (i + newrep > alphabet) { st = BRO_ST_ParseErrorComplexPrefixCodeLengths; break; }
                count = newrep;
            }
#if defined(BRO_SERIAL)
            bro_lens_push_run(sc, lw, i, count, last_nz);
#else
            for (uint32_t k = lane; k < count; k += BRO_W) sc.lens[i + k] = (uint8_t)last_nz;
#endif
            i += count;
            nz += count;
            total += count * (32768u >> last_nz);
            last_repeat = newrep;
            have_repeat = 1;
            if (total == 32768u) break;
            if (total > 32768u) { st = BRO_ST_CodeLengthsChecksum; break; }
            last_symbol = 16;
        } else {
            uint32_t extra;
            if (!bro_read_bits(in, 3, extra)) { st = BRO_ST_UnexpectedEOF; break; }
            if (last_symbol == 17u && have_repeat) {
                uint32_t newrep = 8u * (last_repeat - 2u) + extra + 3u;
                i += newrep - last_repeat;
                last_repeat = newrep;
            } else {
                i += 3u + extra;
                last_repeat = 3u + extra;
            }
            have_repeat = 1;
            if (i > alphabet) { st = BRO_ST_ParseErrorComplexPrefixCodeLengths; break; }
            last_symbol = 17;
        }
    }
#if defined(BRO_SERIAL)
    bro_lens_end(sc, lw);
#endif
    if (st) return st;
    if (nz < 2u) return BRO_ST_LessThanTwoNonZeroCodeLengths;
    bro_syncwarp();
    return 0;
}

// src/lib.rs:877-889.  -> the arguments of the table build that follows: n pairs, explicit symbols or not.
BRO_FN int bro_read_prefix_code_body(BroBits& in, BroScratch& sc, uint32_t alphabet, uint32_t& n_out, bool& explicit_out) {
    uint32_t kind;
    if (!bro_read_bits(in, 2, kind)) return BRO_ST_UnexpectedEOF;
    explicit_out = kind == 1u;
    n_out = alphabet;
    if (kind == 1u) return bro_read_simple_code(in, sc, alphabet, n_out);
    return bro_read_complex_code(in, sc, kind, alphabet);
}

// The out-of-line (cold) routines take the bit window by reference.  Callers hand them a COPY and copy it back, so
// that the hot loops' own window never has its address taken and stays in registers; the routines themselves work on
// a register copy too (BRO_SC_PARAM above).
BRO_COLD int bro_read_prefix_code_cold(BroBits& in_, BRO_SC_PARAM, uint32_t alphabet, uint32_t& n_out, bool& explicit_out) {
    BRO_SC_BIND;
    BroBits in = in_;
    uint32_t n = 0;
    bool ex = false;
    const int st = bro_read_prefix_code_body(in, sc, alphabet, n, ex);
    in_ = in; n_out = n; explicit_out = ex;
    return st;
}

BRO_FN int bro_read_prefix_code(BroBits& in, BroScratch& sc, uint32_t alphabet, uint16_t* T, bool want_root = true) {
    BroBits t = in;
    uint32_t n = 0;
    bool explicit_syms = false;
    int st = bro_read_prefix_code_cold(t, BRO_SC_PASS(sc), alphabet, n, explicit_syms);
    in = t;
    if (st == 0) bro_build_tree(T, BRO_SC_PASS(sc), n, explicit_syms, want_root);
    return st;
}

// src/lib.rs:957-987
BRO_FN int bro_read_block_count(BroBits& in, const uint16_t* T, uint32_t& count) {
    uint32_t sym, extra;
    int r = bro_decode_sym(in, T, sym);
    if (r != BRO_SYM_OK) return BRO_ST_UnexpectedEOF;          // Ok(None) and Err(_) both map to UnexpectedEOF
    if (sym > 25u) return BRO_ST_InvalidBlockCountCode;
    uint32_t be = bro_block_count[sym];
    if (!bro_read_bits(in, be >> 16, extra)) return BRO_ST_UnexpectedEOF;
    count = (be & 0xffffu) + extra;
    return 0;
}

// src/lib.rs:1226-1250 plus the caller's bookkeeping (e.g. 1296-1302)
BRO_FN int bro_block_switch_body(BroBits& in, const uint16_t* arena, BroBlockCat& c) {
    uint32_t code, count;
    int r = bro_decode_sym(in, arena + c.t_type, code);
    if (r == BRO_SYM_HOLE) return BRO_ST_InvalidBlockSwitchCommandCode;
    if (r == BRO_SYM_EOF) return BRO_ST_UnexpectedEOF;
    uint32_t bt = code == 0u ? c.btype_prev : code == 1u ? (c.btype + 1u) % c.nbl : code - 2u;
    int st = bro_read_block_count(in, arena + c.t_count, count);
    if (st) return st;
    c.btype_prev = c.btype;
    c.btype = bt;
    c.blen = count - 1u;
    return 0;
}
BRO_COLD int bro_block_switch_cold(BroBits& in_, const uint16_t* arena, BroBlockCat& c_) {
    BroBits in = in_;
    BroBlockCat c = c_;
    const int st = bro_block_switch_body(in, arena, c);
    in_ = in; c_ = c;
    return st;
}

BRO_FN int bro_block_switch(BroBits& in, const uint16_t* arena, BroBlockCat& c) {
    BroBits t = in;
    BroBlockCat tc = c;
    int st = bro_block_switch_cold(t, arena, tc);
    in = t;
    c = tc;
    return st;
}

// src/lib.rs:1070-1144 and the IMTF of 1164-1177
BRO_FN int bro_read_context_map_body(BroBits& in, BroScratch& sc, uint16_t* T, uint32_t t_cap, uint32_t ntrees, uint32_t len, uint8_t* cmap) {
    const unsigned lane = bro_lane();
    uint32_t b, rlemax = 0;
    if (!bro_read_bits(in, 1, b)) return BRO_ST_UnexpectedEOF;
    if (b) {
        if (!bro_read_bits(in, 4, rlemax)) return BRO_ST_UnexpectedEOF;
        rlemax += 1;
    }
    if (BRO_TREE_U16(rlemax + ntrees) > t_cap) return BRO_ST_ArenaTooSmall;   // temporary table above the arena top
    uint32_t n_pairs = 0;
    bool explicit_syms = false;
    int st = bro_read_prefix_code_body(in, sc, rlemax + ntrees, n_pairs, explicit_syms);
    if (st) return st;
    bro_build_tree(T, BRO_SC_PASS(sc), n_pairs, explicit_syms);
    uint32_t pushed = 0;
    while (pushed < len) {
        uint32_t s;
        int r = bro_decode_sym(in, T, s);
        if (r == BRO_SYM_EOF) return BRO_ST_UnexpectedEOF;
        if (r == BRO_SYM_HOLE) return BRO_ST_ParseErrorContextMap;
        if (s > 0u && s <= rlemax) {
            uint32_t extra;
            if (!bro_read_bits(in, s, extra)) return BRO_ST_UnexpectedEOF;
            uint32_t repeat = (1u << s) + extra;
            if (pushed + repeat > len) return BRO_ST_RunLengthExceededSizeOfContextMap;
            for (uint32_t k = lane; k < repeat; k += BRO_W) cmap[pushed + k] = 0;
            pushed += repeat;
        } else {
            if (lane == 0) cmap[pushed] = (uint8_t)(s == 0u ? 0u : s - rlemax);
            pushed += 1;
        }
    }
    if (!bro_read_bits(in, 1, b)) return BRO_ST_UnexpectedEOF;
    if (b) {
        // 256-entry move-to-front list (the code lengths are dead: their bytes are reused)
#if defined(BRO_SERIAL)
        const BroTlArray<uint8_t, BRO_TL_LENS>& mtf = sc.mtf;
#else
        uint8_t* mtf = sc.lens;
#endif
        bro_syncwarp();
        for (uint32_t k = lane; k < 256u; k += BRO_W) mtf[k] = (uint8_t)k;
        bro_syncwarp();
        if (lane == 0) {
            for (uint32_t k = 0; k < len; k++) {
                uint32_t index = cmap[k];
                uint8_t value = mtf[index];
                cmap[k] = value;
                for (uint32_t j = index; j >= 1u; j--) mtf[j] = (uint8_t)mtf[j - 1];
                mtf[0] = value;
            }
        }
    }
    bro_syncwarp();
    return 0;
}
BRO_COLD int bro_read_context_map_cold(BroBits& in_, BRO_SC_PARAM, uint16_t* T, uint32_t t_cap, uint32_t ntrees, uint32_t len, uint8_t* cmap) {
    BRO_SC_BIND;
    BroBits in = in_;
    const int st = bro_read_context_map_body(in, sc, T, t_cap, ntrees, len, cmap);
    in_ = in;
    return st;
}

BRO_FN int bro_read_context_map(BroBits& in, BroScratch& sc, uint16_t* T, uint32_t t_cap, uint32_t ntrees, uint32_t len, uint8_t* cmap) {
    BroBits t = in;
    int st = bro_read_context_map_cold(t, BRO_SC_PASS(sc), T, t_cap, ntrees, len, cmap);
    in = t;
    return st;
}

// ------------------------------------------------------------------------------------------------------
// byte movement ("phase two"): every routine is executed by the whole warp
// ------------------------------------------------------------------------------------------------------

// 16 bytes from an arbitrary address, as four aligned 32-bit loads + a fifth when misaligned
struct BroV4 { uint32_t x, y, z, w; };
BRO_FN BroV4 bro_load16(const uint8_t* p) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    unsigned sh = 8u * (unsigned)(a & 3u);
    uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    BroV4 r;
    if (sh == 0u) { r.x = w0; r.y = w1; r.z = w2; r.w = w3; }
    else {
        uint32_t w4 = w[4];
        r.x = bro_funnel_r(w0, w1, sh); r.y = bro_funnel_r(w1, w2, sh);
        r.z = bro_funnel_r(w2, w3, sh); r.w = bro_funnel_r(w3, w4, sh);
    }
    return r;
}

// dst[0..n) = src[0..n) where the regions do not overlap, or dst - src >= 16*BRO_W (so that one warp step never
// reads a byte written in the same step).  dst is advanced in 16-byte aligned vector stores.
BRO_COPY_FN void bro_copy_far(uint8_t* dst, const uint8_t* src, uint32_t n) {
    const unsigned lane = bro_lane();
#if defined(BRO_HOSTSIM)
    for (uint32_t i = 0; i < n; i++) dst[i] = src[i];
    (void)lane;
#elif defined(BRO_THREAD_MODE)
    // one thread: byte head until dst is 16-byte aligned, 16-byte stores, byte tail.  Overlap is fine when
    // dst - src >= 16 (a vector step only reads bytes this thread wrote in earlier steps).
    (void)lane;
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    for (uint32_t i = 0; i < head; i++) dst[i] = src[i];
    dst += head; src += head; n -= head;
    for (uint32_t k = n >> 4; k; k--) {
        BroV4 v = bro_load16(src);
        *(uint4*)dst = make_uint4(v.x, v.y, v.z, v.w);
        dst += 16; src += 16;
    }
    n &= 15u;
    for (uint32_t i = 0; i < n; i++) dst[i] = src[i];
#else
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    if (lane < head) dst[lane] = src[lane];
    if (BRO_W < 16u && lane + BRO_W < head) dst[lane + BRO_W] = src[lane + BRO_W];
    dst += head; src += head; n -= head;
    while (n >= 16u * BRO_W) {
        bro_syncwarp();
        BroV4 v = bro_load16(src + 16u * lane);
        *(uint4*)(dst + 16u * lane) = make_uint4(v.x, v.y, v.z, v.w);
        dst += 16u * BRO_W; src += 16u * BRO_W; n -= 16u * BRO_W;
    }
    bro_syncwarp();
    uint32_t nv = n >> 4;
    if (lane < nv) {
        BroV4 v = bro_load16(src + 16u * lane);
        *(uint4*)(dst + 16u * lane) = make_uint4(v.x, v.y, v.z, v.w);
    }
    uint32_t done = nv << 4;
    if (done + lane < n) dst[done + lane] = src[done + lane];
    if (BRO_W < 16u && done + lane + BRO_W < n) dst[done + lane + BRO_W] = src[done + lane + BRO_W];
#endif
}

// LZ77 backward copy inside the output slot: out[pos .. pos+len) = bytes `dist` back, periodic when dist < len
// (src/lib.rs:1491-1505; the ring-buffer window of src/ringbuffer/mod.rs is the linear output itself).
BRO_COPY_FN void bro_lz_copy(uint8_t* out, uint32_t pos, uint32_t dist, uint32_t len) {
    const unsigned lane = bro_lane();
    bro_syncwarp();
    uint8_t* dst = out + pos;
    const uint8_t* src = dst - dist;
#if defined(BRO_HOSTSIM)
    for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
    (void)lane;
#elif defined(BRO_THREAD_MODE)
    (void)lane;
    if (len < 16u) { for (uint32_t i = 0; i < len; i++) dst[i] = src[i]; return; }
    if (dist >= 16u) { bro_copy_far(dst, src, len); return; }
    // short period: bytes until m*dist >= 16 bytes of the pattern exist, then 16-byte steps from m*dist back
    uint32_t m = (15u + dist) / dist;
    uint32_t n0 = (m - 1u) * dist;
    if (n0 > len) n0 = len;
    for (uint32_t i = 0; i < n0; i++) dst[i] = src[i];
    if (len > n0) bro_copy_far(dst + n0, dst + n0 - m * dist, len - n0);
#else
    if (len <= BRO_W) {
        if (lane < len) dst[lane] = src[dist >= len ? lane : lane % dist];
        return;
    }
    if (dist >= len || dist >= 16u * BRO_W) { bro_copy_far(dst, src, len); return; }
    // periodic fill: the first (m-1)*dist bytes straight from the pattern, the rest from m*dist >= 16*BRO_W bytes back
    uint32_t m = (16u * BRO_W - 1u + dist) / dist;
    uint32_t n0 = (m - 1u) * dist;
    if (n0 > len) n0 = len;
    for (uint32_t i = lane; i < n0; i += BRO_W) dst[i] = src[i % dist];
    // the lanes leave the loop above at different trips, and the head bytes of the copy below read what OTHER lanes have
    // just written (m * dist >= 16 * BRO_W bytes back is inside the prefix): order the stores before those loads
    bro_syncwarp();
    if (len > n0) bro_copy_far(dst + n0, dst + n0 - m * dist, len - n0);
#endif
}

// Static dictionary word + transform (src/lib.rs:1506-1540, src/transformation/mod.rs:84-209).  Returns the
// transformed length, or -1 where the reference panics (uppercase_first on a 0x00 byte).
BRO_COLD int bro_dict_word(BRO_SC_PARAM, const uint8_t* dict, int quirk_spec, uint32_t copy_len, uint32_t index, uint32_t tid) {
    BRO_SC_BIND;
    const unsigned lane = bro_lane();
    const uint8_t* w = dict + bro_dict_offsets[copy_len] + index * copy_len;
    uint32_t type = bro_xf_type[tid], plen = bro_xf_prefix_len[tid], slen = bro_xf_suffix_len[tid];
    uint32_t from = 0, wl = copy_len;
    if (type >= 3u && type <= 11u) {            // OmitFirstN: base_word[min(N, len-1)..] (Q3) / spec: [min(N,len)..]
        uint32_t n = type - 2u;
        from = quirk_spec ? (n < wl ? n : wl) : (n < wl - 1u ? n : wl - 1u);
        wl -= from;
    } else if (type >= 12u) {                   // OmitLastN: base_word[..max(N,len)-N]
        uint32_t n = type - 11u;
        wl = (wl > n ? wl : n) - n;
    }
    bro_syncwarp();
    for (uint32_t i = lane; i < plen; i += BRO_W) sc.word[i] = bro_xf_strings[bro_xf_prefix_off[tid] + i];
    for (uint32_t i = lane; i < wl; i += BRO_W) sc.word[plen + i] = w[from + i];
    for (uint32_t i = lane; i < slen; i += BRO_W) sc.word[plen + wl + i] = bro_xf_strings[bro_xf_suffix_off[tid] + i];
    bro_syncwarp();
    int ret = (int)(plen + wl + slen);
    if (type == 1u || type == 2u) {
        // uppercase_first (src/transformation/mod.rs:42-82) / uppercase_all (3-40): a serial UTF-8 walk
        uint32_t c0 = sc.word[plen];
        if (type == 1u && c0 == 0u && !quirk_spec) ret = -1;
        else if (lane == 0) {
            uint32_t i = 0;                     // position in the transformed word, which starts at sc.word[plen]
            do {
                uint32_t c = sc.word[plen + i];
                if (c < 192u) { if (c >= 97u && c <= 122u) sc.word[plen + i] = (uint8_t)(c ^ 32u); i += 1; }
                else if (c < 224u) { if (i + 1 < wl) sc.word[plen + i + 1] = (uint8_t)((uint32_t)sc.word[plen + i + 1] ^ 32u); i += 2; }
                else { if (i + 2 < wl) sc.word[plen + i + 2] = (uint8_t)((uint32_t)sc.word[plen + i + 2] ^ 5u); i += 3; }
            } while (type == 2u && i < wl);
        }
        bro_syncwarp();
    }
    return ret;
}

// "Phase two" of a command: materialise the copy -- an LZ77 back-reference into the output produced so far, or a
// static dictionary word with its transform (src/lib.rs:1483-1542 and 2102-2124).  track_ctx: refresh the literal
// context bytes p1/p2 from the output (not needed while a meta-block has no context modelling).
BRO_FN int bro_emit_copy(BroDec& d, BroScratch& sc, uint32_t mlen, uint32_t mb_begin, uint32_t distance, uint32_t max_allowed,
                         uint32_t copy_len, bool track_ctx) {
    const unsigned lane = bro_lane();
    uint32_t mb_out = d.pos - mb_begin;
    if (distance <= max_allowed) {
        if (mlen < mb_out + copy_len) return BRO_ST_ExceededExpectedBytes;     // src/lib.rs:2105-2108
        if (copy_len > d.cap - d.pos) return BRO_ST_OutputTooSmall;
#if !defined(BRO_SERIAL) && BRO_W == 32u
        if (track_ctx && copy_len <= BRO_W) {
            // the copy of a text-like stream (8 bytes on average): a byte per lane, in line, and the two bytes the next literal's
            // context is made of come from the lanes that moved them instead of a read-back from memory
            bro_syncwarp();
            uint8_t* dst = d.out + d.pos;
            const uint8_t* src = dst - distance;
            uint32_t idx = lane, v = 0;
            if (distance < copy_len) idx = lane % distance;
            if (lane < copy_len) { v = src[idx]; dst[lane] = (uint8_t)v; }
            d.pos += copy_len;
            d.p1 = bro_shfl(v, copy_len - 1u); d.p2 = bro_shfl(v, copy_len - 2u);                          // copy_len >= 2
            return 0;
        }
#endif
        bro_lz_copy(d.out, d.pos, distance, copy_len);
        d.pos += copy_len;
        if (track_ctx) {
            bro_syncwarp();
            d.p1 = d.out[d.pos - 1]; d.p2 = d.out[d.pos - 2];                    // copy_len >= 2
        }
    } else {
        if (copy_len < 4u || copy_len > 24u) return BRO_ST_InvalidLengthInStaticDictionary;
        uint32_t word_id = distance - max_allowed - 1u;
        uint32_t bits = bro_dict_size_bits[copy_len];
        uint32_t index = word_id & ((1u << bits) - 1u), tid = word_id >> bits;
        if (tid > 120u) return BRO_ST_InvalidTransformId;
        int n = bro_dict_word(BRO_SC_PASS(sc), d.dict, d.quirk_spec, copy_len, index, tid);
        if (n < 0) return BRO_ST_PanicUppercaseZero;
        if (mlen < mb_out + (uint32_t)n) return BRO_ST_ExceededExpectedBytes;  // checked after the transform (Q10)
        if ((uint32_t)n > d.cap - d.pos) return BRO_ST_OutputTooSmall;
        for (uint32_t i = lane; i < (uint32_t)n; i += BRO_W) d.out[d.pos + i] = sc.word[i];
        if (n >= 2) { d.p1 = sc.word[n - 1]; d.p2 = sc.word[n - 2]; }
        else if (n == 1) { d.p2 = d.p1; d.p1 = sc.word[0]; }
        d.pos += (uint32_t)n;
        bro_syncwarp();
    }
    return 0;
}

// src/lib.rs:1412-1481: distance from the distance code (ring-buffer codes 0-15, direct codes, or extra bits), and
// the ring-buffer update rule.  Returns 0 or a status.
BRO_FN int bro_resolve_distance(BroDec& d, uint32_t dcode, uint32_t npostfix, uint32_t ndirect, uint32_t& distance,
                                uint32_t& max_allowed) {
    if (dcode <= 3u) distance = dcode == 0u ? d.d0 : dcode == 1u ? d.d1 : dcode == 2u ? d.d2 : d.d3;
    else if (dcode <= 15u) {
        int32_t basev = (int32_t)(dcode <= 9u ? d.d0 : d.d1);
        int32_t delta = (int32_t)(((dcode <= 9u ? dcode - 2u : dcode - 8u)) >> 1);
        int64_t v = (int64_t)(uint32_t)basev + ((dcode & 1u) ? (int64_t)delta : -(int64_t)delta);
        if (v <= 0) return BRO_ST_InvalidNonPositiveDistance;
        distance = (uint32_t)v;
    } else if (dcode <= 15u + ndirect) distance = dcode - 15u;
    else {
        uint32_t t = dcode - ndirect - 16u;
        uint32_t ndistbits = 1u + (t >> (npostfix + 1u));
        uint32_t dextra;
        if (!bro_read_bits(d.in, ndistbits, dextra)) return BRO_ST_UnexpectedEOF;
        uint32_t hcode = t >> npostfix, lcode = t & ((1u << npostfix) - 1u);
        uint32_t offset = ((2u + (hcode & 1u)) << ndistbits) - 4u;
        distance = ((offset + dextra) << npostfix) + lcode + ndirect + 1u;
    }
    max_allowed = d.window < d.pos ? d.window : d.pos;
    if (dcode > 0u && distance <= max_allowed) {                                // src/lib.rs:1476-1478
        d.d3 = d.d2; d.d2 = d.d1; d.d1 = d.d0; d.d0 = distance;
    }
    return 0;
}

#if !defined(BRO_PARSE)   /* the fused command loops (the two-phase path has its own: bro_parse.h) */
// Reference to a 256-entry root table held on chip (shared memory in the group modes): a 32-bit shared-window
// address, so that a lookup is one add and one LDS and the base lives in ONE register for the whole loop.
#if defined(BRO_SERIAL) || defined(BRO_WARPSIM)   /* (BRO_WARPSIM: the 32-lane form compiled for the host, bro_warpsim.cpp -- CPU test-suite only) */
typedef const uint16_t* BroRoot;
BRO_FN BroRoot bro_root_ref(const uint16_t* p) { return p; }
BRO_FN uint32_t bro_root_get(BroRoot r, uint32_t idx) { return r[idx]; }
#else
typedef uint32_t BroRoot;
BRO_FN BroRoot bro_root_ref(const uint16_t* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a) :: "memory");    // opaque: keep it in a register instead of recomputing it per symbol
    return a;
}
BRO_FN uint32_t bro_root_get(BroRoot r, uint32_t idx) {
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(r + 2u * idx));
    return v;
}
#endif

// One symbol through an on-chip root table.  No end-of-input test on the fast path: the window's `avail` simply goes
// negative, which is sticky (nothing real is ever loaded again), and the callers test it at their checkpoints --
// before any decision that could turn garbage bits into a different error, and before any copy is materialised.
// Returns BRO_SYM_OK or, from the slow path, BRO_SYM_EOF / BRO_SYM_HOLE.
BRO_FN int bro_decode_sym_onchip(BroBits& s, BroRoot root, const uint16_t* T, uint32_t& sym) {
    bro_refill(s);
    uint32_t peek = bro_peek(s);
    uint32_t e = bro_root_get(root, peek & 0xffu);
    uint32_t len = e >> 10;
    if (len != 0u) {
        bro_consume(s, len);
        sym = e & 0x3ffu;
        return BRO_SYM_OK;
    }
    uint32_t r = bro_sym_slow(T, peek, e, (int32_t)s.avail < 0 ? 0u : s.avail);
    bro_consume(s, (r >> 16) & 0xffu);
    sym = r & 0xffffu;
    return (int)(r >> 24);
}

// The command loop of a "simple" meta-block: one literal, one insert&copy and one distance code, no block
// switches -- i.e. no context modelling, which is every meta-block libbrotli emits at quality <= 9 and the case
// BASELINE's high-ratio workload consists of.  Same results as the general loop below (src/lib.rs:2003-2141), but the
// three root tables are on chip, literals are not tracked as context, and end-of-input is tested at checkpoints.
BRO_FN int bro_commands_simple(BroDec& d, BroScratch& sc, uint32_t mlen, uint32_t npostfix, uint32_t ndirect,
                               const uint16_t* T_lit, const uint16_t* T_cmd, const uint16_t* T_dist) {
    const unsigned lane = bro_lane();
    const uint32_t mb_begin = d.pos;
    const BroRoot r_lit = bro_root_ref(sc.root_lit), r_cmd = bro_root_ref(sc.root_cmd), r_dist = bro_root_ref(sc.root_dist);
    int st = 0;
    for (;;) {
        // ---- phase one: entropy decode of one insert&copy command ----
        uint32_t sym, v;
        int r = bro_decode_sym_onchip(d.in, r_cmd, T_cmd, sym);
        if (r != BRO_SYM_OK) { st = r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertAndCopyLength : BRO_ST_UnexpectedEOF; break; }
        const uint32_t ie = bro_ic_insert[sym], ce = bro_ic_copy[sym];
        uint32_t insert_len = ie & 0xffffu, copy_len = ce & 0xffffu;
        if (!bro_read_bits(d.in, ie >> 16, v)) { st = BRO_ST_UnexpectedEOF; break; }     // insert extra bits first
        insert_len += v;
        if (!bro_read_bits(d.in, ce >> 16, v)) { st = BRO_ST_UnexpectedEOF; break; }
        copy_len += v;
        if ((int32_t)d.in.avail < 0) { st = BRO_ST_UnexpectedEOF; break; }                // checkpoint
        if (mlen < (d.pos - mb_begin) + insert_len) { st = BRO_ST_ExceededExpectedBytes; break; }   // src/lib.rs:2036-2039
        // literals: lane k%W keeps literal k, the group stores W literals with one coalesced store.  A run that does
        // not fit the slot is still decoded (a decode error wins over OutputTooSmall, as in the general loop).
        const bool fits = insert_len <= d.cap - d.pos;
        uint8_t* o = d.out + d.pos;
        uint32_t k = 0;
#if !defined(BRO_SERIAL) && BRO_W == 32u
        // LANE-PARALLEL literal decode.  All literals of the run use one code, so lane l looks up the code that would
        // start at bit offset l of the next 32 stream bits; the symbols actually present are the chain 0 -> len(0) ->
        // len(0)+len(len(0)) ... which is walked with one shuffle per symbol, and every chain member stores its
        // literal at its rank.  A round therefore costs a handful of instructions per literal instead of a full
        // serial decode; codes longer than 8 bits end a round and take the scalar path.
        while (k < insert_len) {
            bro_refill(d.in);
            const uint32_t w2 = bro_shfl(d.in.cur, d.in.wi);                  // the word after the window, not consumed
            const uint32_t lo = bro_funnel_r(d.in.w0, d.in.w1, d.in.bp);     // stream bits [0, 32)
            const uint32_t hi = bro_funnel_r(d.in.w1, w2, d.in.bp);          // stream bits [32, 64)
            const uint32_t e = bro_root_get(r_lit, bro_funnel_r(lo, hi, lane) & 0xffu);
            const uint32_t len_l = e >> 10;                                   // 0: longer than 8 bits (or a hole)
            uint32_t want = insert_len - k, pos = 0, mask = 0, cnt = 0;
            while (cnt < want) {
                const uint32_t L = bro_shfl(len_l, pos);
                if (L == 0u || pos + L > 32u) break;
                mask |= 1u << pos;
                pos += L;
                cnt++;
                if (pos >= 32u) break;
            }
            if (cnt != 0u) {
                if (fits && ((mask >> lane) & 1u)) o[k + bro_popc(mask & bro_lanemask_lt())] = (uint8_t)e;
                k += cnt;
                bro_consume(d.in, pos);
            } else {
                uint32_t lit;
                r = bro_decode_sym_onchip(d.in, r_lit, T_lit, lit);
                if (r != BRO_SYM_OK) { st = r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF; break; }
                if (fits && lane == 0) o[k] = (uint8_t)lit;
                k++;
            }
        }
        if (st) break;
        if ((int32_t)d.in.avail < 0) { st = BRO_ST_UnexpectedEOF; break; }                // checkpoint
        if (!fits) { d.pos = d.cap; st = BRO_ST_OutputTooSmall; break; }
#else
        uint32_t mine = 0;
        while (k < insert_len) {
            uint32_t lit;
            r = bro_decode_sym_onchip(d.in, r_lit, T_lit, lit);
            if (r != BRO_SYM_OK) { st = r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF; break; }
            if (lane == (k & (BRO_W - 1u))) mine = lit;
            k++;
            if ((k & (BRO_W - 1u)) == 0u && fits) o[k - BRO_W + lane] = (uint8_t)mine;
        }
        if (st) break;
        if ((int32_t)d.in.avail < 0) { st = BRO_ST_UnexpectedEOF; break; }                // checkpoint
        if (!fits) { d.pos = d.cap; st = BRO_ST_OutputTooSmall; break; }
        {
            uint32_t tail = k & (BRO_W - 1u);
            if (lane < tail) o[k - tail + lane] = (uint8_t)mine;
        }
#endif
        d.pos += insert_len;
        if (d.pos - mb_begin == mlen) break;                                              // src/lib.rs:2069-2070
        // distance code (src/lib.rs:1367-1410) and distance (1412-1481)
        uint32_t dcode = 0;
        if (sym >= 128u) {
            r = bro_decode_sym_onchip(d.in, r_dist, T_dist, dcode);
            if (r != BRO_SYM_OK) { st = r == BRO_SYM_HOLE ? BRO_ST_ParseErrorDistanceCode : BRO_ST_UnexpectedEOF; break; }
        }
        uint32_t distance, max_allowed;
        if ((st = bro_resolve_distance(d, dcode, npostfix, ndirect, distance, max_allowed))) {
            if ((int32_t)d.in.avail < 0) st = BRO_ST_UnexpectedEOF;
            break;
        }
        if ((int32_t)d.in.avail < 0) { st = BRO_ST_UnexpectedEOF; break; }                // checkpoint
        // ---- phase two: materialise the copy ----
        if ((st = bro_emit_copy(d, sc, mlen, mb_begin, distance, max_allowed, copy_len, false))) break;
        if (d.pos - mb_begin == mlen) break;                                              // src/lib.rs:2128-2130
    }
    if (st == 0) {
        // literal context for the following meta-block (literal_buf persists)
        bro_syncwarp();
        uint32_t m = d.pos - mb_begin;
        if (m >= 2u) { d.p1 = d.out[d.pos - 1]; d.p2 = d.out[d.pos - 2]; }
        else { d.p2 = d.p1; d.p1 = d.out[d.pos - 1]; }
    }
    return st;
}

#endif

// ------------------------------------------------------------------------------------------------------
// one compressed meta-block: src/lib.rs:1745-2141
// ------------------------------------------------------------------------------------------------------
BRO_FN int bro_step_block(BroDec& d, BroBlockCat& c) {   // src/lib.rs:1182-1197 et al.
    if (c.nbl < 2u) return 0;
    if (c.blen == 0u) return bro_block_switch(d.in, d.arena, c);
    c.blen -= 1u;
    return 0;
}

// What the header of a compressed meta-block announces (src/lib.rs:1745-2002), with the arena offsets of its tables.
struct BroMbInfo {
    BroBlockCat cat[3];
    uint32_t npostfix, ndirect, ntl, ntd;
    uint32_t o_modes, o_cmap_l, o_cmap_d, o_lit, o_cmd, o_dist, dist_stride;
    bool simple;              // one literal / insert&copy / distance code and no block switches: no context modelling
    bool lctx;                // (two-phase path) the literal code of a symbol depends on the two bytes in front of it
};

// Header of a compressed meta-block: block-type codes, NPOSTFIX/NDIRECT, context modes and maps, and all prefix code
// tables, bump-allocated in the arena.
BRO_FN int bro_metablock_tables(BroDec& d, BroMbInfo& mb) {
    const unsigned lane = bro_lane();
    uint16_t* A = d.arena;
    BroBlockCat (&cat)[3] = mb.cat;
    BroScratch& sc = BRO_DSC(d);
    // Single exit: an error only sets `st` and the remaining steps are skipped, so that in the thread-per-stream kernel
    // the lanes of a warp meet again after every conditional part (a `return` inside would let the lanes that skip a
    // part run ahead of the others).
    int st = 0;
    // The arena is bump-allocated per meta-block from the counts the header announces (uint16 units, 16-byte
    // granules); a meta-block that does not fit reports ArenaTooSmall and the stream is re-run by the warp kernel,
    // whose arenas hold the worst case (256 + 256 + 256 tables).
    uint32_t top = d.arena_base;
#define BRO_ALLOC(var, n_u16) do { (var) = top; top += ((uint32_t)(n_u16) + 7u) & ~7u; if (!st && top > d.arena_cap) st = BRO_ST_ArenaTooSmall; } while (0)
#define BRO_TRY(expr) do { if (!st) st = (expr); } while (0)
    // NBLTYPES{L,I,D}, block type / count codes, first block counts (src/lib.rs:1745-1885)
#pragma unroll
    for (uint32_t k = 0; k < 3u; k++) {
        cat[k].btype = 0; cat[k].btype_prev = 1; cat[k].blen = 0; cat[k].t_type = 0; cat[k].t_count = 0; cat[k].nbl = 1;
        BRO_TRY(bro_read_nbltypes(d.in, cat[k].nbl));
        if (!st && cat[k].nbl >= 2u) {
            BRO_ALLOC(cat[k].t_type, BRO_TREE_U16(cat[k].nbl + 2u));
            BRO_ALLOC(cat[k].t_count, BRO_TREE_U16(BRO_ALPHA_BCOUNT));
            // the two codes of a category are read by one loop so that the table reader has a single call site here
            for (uint32_t j = 0; j < 2u; j++)
                BRO_TRY(bro_read_prefix_code(d.in, sc, j == 0u ? cat[k].nbl + 2u : BRO_ALPHA_BCOUNT,
                                             A + (j == 0u ? cat[k].t_type : cat[k].t_count)));
            BRO_TRY(bro_read_block_count(d.in, A + cat[k].t_count, cat[k].blen));
        }
    }
    // NPOSTFIX, NDIRECT (src/lib.rs:548-560), context modes (562-573)
    uint32_t npostfix = 0, ndirect = 0;
    if (!st && !bro_read_bits(d.in, 2, npostfix)) st = BRO_ST_UnexpectedEOF;
    if (!st && !bro_read_bits(d.in, 4, ndirect)) st = BRO_ST_UnexpectedEOF;
    ndirect <<= npostfix;
    uint32_t o_modes = 0, o_cmap_l = 0, o_cmap_d = 0;
    BRO_ALLOC(o_modes, (cat[0].nbl + 1u) >> 1);
    uint8_t* modes = (uint8_t*)(A + o_modes);
    for (uint32_t i = 0; !st && i < cat[0].nbl; i++) {
        uint32_t m = 0;
        if (!bro_read_bits(d.in, 2, m)) st = BRO_ST_UnexpectedEOF;
        else if (lane == 0) modes[i] = (uint8_t)m;
    }
    // NTREESL + literal context map, NTREESD + distance context map (src/lib.rs:1916-1973)
    BRO_ALLOC(o_cmap_l, 32u * cat[0].nbl);
    BRO_ALLOC(o_cmap_d, 2u * cat[2].nbl);
    uint8_t* cmap_l = (uint8_t*)(A + o_cmap_l);
    uint8_t* cmap_d = (uint8_t*)(A + o_cmap_d);
    uint32_t ntl = 1, ntd = 1;
    bool lctx = false;
    for (uint32_t j = 0; j < 2u; j++) {
        uint32_t nt = 1;
        BRO_TRY(bro_read_nbltypes(d.in, nt));
        if (!st && nt >= 2u) {
            // the context map's own prefix code is temporary: it lives above the arena top and is dropped afterwards
            st = bro_read_context_map(d.in, sc, A + top, d.arena_cap - top, nt, j == 0u ? 64u * cat[0].nbl : 4u * cat[2].nbl,
                                      j == 0u ? cmap_l : cmap_d);
        }
        if (j == 0u) ntl = nt; else ntd = nt;
#if defined(BRO_PARSE)
        // Phase one of the two-phase path never sees the bytes copies produce, so it can decode a meta-block only if
        // no block type's literal context map depends on the context; found out here, before the prefix codes
        // (the bulk of the header) are read.
        if (!st && j == 0u && nt >= 2u) {
            bool dep = false;
            for (uint32_t q = 0; q < 64u * cat[0].nbl; q++)
                if (cmap_l[q] != cmap_l[q & ~63u]) dep = true;
            if (dep) {
                // ... unless the thread executes the stream's copies itself (bro_parse.h, immediate mode): possible when
                // nothing of the stream has been left to phase two yet, and not when the stream is only measured
                if (d.sizing || (!d.imm && d.nrec != 0u)) st = BRO_ST_NeedFused;
                else { d.imm = 1u; lctx = true; }
            }
        }
#endif
    }
    // prefix codes (src/lib.rs:1016-1068): NTREESL literal codes, NBLTYPESI insert&copy codes, NTREESD distance codes
    const uint32_t dist_alphabet = 16u + ndirect + (48u << npostfix);
    const uint32_t dist_stride = BRO_TREE_U16(dist_alphabet);
    uint32_t o_lit = 0, o_cmd = 0, o_dist = 0;
#if defined(BRO_PARSE)
    top = (top + 63u) & ~63u;
#endif
    BRO_ALLOC(o_lit, ntl * BRO_LIT_STRIDE_U16);
    BRO_ALLOC(o_cmd, cat[1].nbl * BRO_TREE_U16(BRO_ALPHA_CMD));
    BRO_ALLOC(o_dist, ntd * dist_stride);
#undef BRO_ALLOC
    {
        const uint32_t n_l = ntl, n_i = cat[1].nbl, total = ntl + cat[1].nbl + ntd;
        for (uint32_t i = 0; !st && i < total; i++) {
            uint32_t alphabet;
            uint16_t* T;
            if (i < n_l) { alphabet = BRO_ALPHA_LIT; T = A + o_lit + i * BRO_LIT_STRIDE_U16; }
            else if (i < n_l + n_i) { alphabet = BRO_ALPHA_CMD; T = A + o_cmd + (i - n_l) * BRO_TREE_U16(BRO_ALPHA_CMD); }
            else { alphabet = dist_alphabet; T = A + o_dist + (i - n_l - n_i) * dist_stride; }
#if defined(BRO_PARSE)
            // literal and insert&copy tables, and the distance table when there is one, are decoded canonically
            // (bro_parse.h): no root.  Literal codes chosen per context are looked up in the arena: root.
            st = bro_read_prefix_code(d.in, sc, alphabet, T, i < n_l ? lctx : (i >= n_l + n_i && ntd >= 2u));
#else
            st = bro_read_prefix_code(d.in, sc, alphabet, T);
#endif
        }
    }
#undef BRO_TRY
    bro_syncwarp();
#if defined(BRO_HOSTSIM) && defined(BRO_MB_STAT)
    BRO_MB_STAT(ntl, ntd, cat[0].nbl, cat[1].nbl, cat[2].nbl, st);
#endif
    mb.npostfix = npostfix; mb.ndirect = ndirect; mb.ntl = ntl; mb.ntd = ntd;
    mb.o_modes = o_modes; mb.o_cmap_l = o_cmap_l; mb.o_cmap_d = o_cmap_d;
    mb.o_lit = o_lit; mb.o_cmd = o_cmd; mb.o_dist = o_dist; mb.dist_stride = dist_stride;
    mb.simple = ntl == 1u && ntd == 1u && cat[0].nbl == 1u && cat[1].nbl == 1u && cat[2].nbl == 1u;
    mb.lctx = lctx;
    return st;
}

#if !defined(BRO_PARSE)
// One symbol through a narrow on-chip copy of a table's root (`hot`: 1 << rb entries, an entry is a direct hit -- symbol |
// len << 10, len <= rb -- or 1 = settle it in the table T itself).  Same results as bro_decode_sym on T.
BRO_FN int bro_decode_sym_hot(BroBits& s, const uint16_t* hot, uint32_t rb, const uint16_t* T, uint32_t& sym) {
    bro_refill(s);
    const uint32_t peek = bro_peek(s);
    uint32_t e = hot[peek & ((1u << rb) - 1u)];
    uint32_t len = e >> 10;
    if (len == 0u) {
        e = T[peek & (BRO_ROOT_SIZE - 1u)];
        len = e >> 10;
        if (len == 0u) {
            const uint32_t r = bro_sym_slow(T, peek, e, bro_avail(s));
            bro_consume(s, (r >> 16) & 0xffu);
            sym = r & 0xffffu;
            return (int)(r >> 24);
        }
    }
    if (len > bro_avail(s)) return BRO_SYM_EOF;
    bro_consume(s, len);
    sym = e & 0x3ffu;
    return BRO_SYM_OK;
}


#if !defined(BRO_SERIAL) && BRO_W == 32u
// The literal code of block type bt when it does not depend on the context -- the 64 entries of the type's context map row
// name one tree (every stream without literal context modelling that has literal block types: libbrotli quality 5..9 on
// mixed data) -- or BRO_ROW_CTX.  The 8-bit root of that tree is staged in sc.root_lit (unless the meta-block has one
// literal code, whose root bro_stage_roots staged), so that the runs of the block are decoded lane-parallel.
#define BRO_ROW_CTX 0xffffffffu
BRO_FN uint32_t bro_lit_row(BroScratch& sc, uint16_t* root10, const uint16_t* T_lit, const uint8_t* cmap_l, uint32_t ntl, uint32_t bt) {
    const unsigned lane = bro_lane();
    uint32_t t0 = 0;
    if (ntl >= 2u) {
        const uint8_t* row = cmap_l + 64u * bt;
        t0 = row[0];
        const bool same = row[lane] == t0 && row[32u + lane] == t0;
        if (!__all_sync(0xffffffffu, same)) return BRO_ROW_CTX;
    }
    bro_syncwarp();                                                       // (lanes may still be reading the roots staged before)
    const uint16_t* T = T_lit + t0 * BRO_LIT_STRIDE_U16;
    if (ntl >= 2u) for (uint32_t r = lane; r < 256u; r += BRO_W) sc.root_lit[r] = T[r];
    if (root10 == 0) { bro_syncwarp(); return t0; }
    // the 10-bit root, from the canonical form of the code: entry i answers the stream bits i (first bit = bit 0).  Real data
    // has literal codes of 9 and 10 bits all the time (data/metablock_reset: a third of its literals), and a code the root
    // does not answer ends a lane-parallel round
    uint32_t lim[10];
    int base[10];
#pragma unroll
    for (uint32_t L = 1; L <= 10u; L++) { lim[L - 1u] = T[BRO_T_LIMIT + L]; base[L - 1u] = (int)(int16_t)T[BRO_T_BASE + L]; }
    const bool single = T[BRO_T_SINGLE] != 0;
    for (uint32_t i = lane; i < 1024u; i += BRO_W) {
        const uint32_t x = bro_brev(i) >> 17;                             // the ten bits left-justified in 15, first bit most significant
        uint32_t e = 0;
#pragma unroll
        for (uint32_t L = 10u; L >= 1u; L--)                              // the shortest length whose limit x stays under
            if (x < lim[L - 1u]) e = (uint32_t)(base[L - 1u] + (int)(x >> (15u - L))) | (L << 16);
        if (e != 0u && !single) e = (uint32_t)T[BRO_T_SORTED + (e & 0xffffu)] | ((e >> 16) << 10);
        else e = 0;
        root10[i] = (uint16_t)e;
    }
    bro_syncwarp();
    return t0;
}
#endif
// What the general loop keeps on chip for a meta-block with several codes of a kind
struct BroHot {
    uint32_t rb_lit, rb_cmd, rb_dist;    // root bits of the on-chip copies (0 = none: look the tables up in the arena)
    bool cmap_l, cmap_d, modes;          // the context maps / modes fit their on-chip copies
};

// The general command loop (src/lib.rs:2003-2141): any number of codes, block switches, literal context modelling.
// Per literal the reference looks up the context mode of the block type, the context map, and then the code the map
// names (src/lib.rs:1317-1345).  In round 1 all three were reads of the warp's arena in HBM / L2, one after the other:
// ~1,400 cycles per symbol on a single warp, which is what bounds a long context-modelled stream (and any batch that
// holds one).  They are shared-memory look-ups now whenever the meta-block's tables fit (bro_stage_hot).
BRO_FN int bro_commands_general(BroDec& d, BroScratch& sc, uint32_t mlen, BroMbInfo& mb, const BroHot& hot) {
    const unsigned lane = bro_lane();
    uint16_t* A = d.arena;
    BroBlockCat (&cat)[3] = mb.cat;
    int st;
    const uint32_t npostfix = mb.npostfix, ndirect = mb.ndirect, ntl = mb.ntl, ntd = mb.ntd, dist_stride = mb.dist_stride;
    const uint8_t* modes = hot.modes ? (const uint8_t*)sc.hot_modes : (const uint8_t*)(A + mb.o_modes);
    const uint8_t* cmap_l = hot.cmap_l ? (const uint8_t*)sc.hot_cmap_l : (const uint8_t*)(A + mb.o_cmap_l);
    const uint8_t* cmap_d = hot.cmap_d ? (const uint8_t*)sc.hot_cmap_d : (const uint8_t*)(A + mb.o_cmap_d);
    const uint16_t* const T_lit = A + mb.o_lit;
    const uint16_t* const T_cmd = A + mb.o_cmd;
    const uint16_t* const T_dist = A + mb.o_dist;
    // one code of a kind: its 8-bit root table is on chip (bro_stage_roots); several: narrow roots of all of them (bro_stage_hot)
    const bool one_lit = ntl == 1u, one_cmd = cat[1].nbl == 1u, one_dist = ntd == 1u;
    const bool lit_simple = one_lit && cat[0].nbl == 1u;   // no context modelling, no literal block switches
    const uint32_t mb_begin = d.pos;   // meta_block.count_output == d.pos - mb_begin
    // (the block categories live in the caller's frame, i.e. in local memory: whether a category has block types at all is asked
    // once, not at every symbol)
    const bool types0 = cat[0].nbl >= 2u, types1 = cat[1].nbl >= 2u, types2 = cat[2].nbl >= 2u;
#if !defined(BRO_SERIAL) && BRO_W == 32u
    uint32_t row = bro_lit_row(sc, d.root10, T_lit, cmap_l, ntl, cat[0].btype);   // the tree of the current literal block type, or BRO_ROW_CTX
    (void)lit_simple;
#endif
    // command loop (src/lib.rs:2003-2141)
    for (;;) {
        // ---- phase one: entropy decode of one insert&copy command ----
        uint32_t sym, extra;
        if (types1 && (st = bro_step_block(d, cat[1]))) return st;
        int r;
        if (one_cmd) r = bro_decode_sym2(d.in, sc.root_cmd, T_cmd, sym);
        else if (hot.rb_cmd) r = bro_decode_sym_hot(d.in, sc.root_cmd + (cat[1].btype << hot.rb_cmd), hot.rb_cmd,
                                                    T_cmd + cat[1].btype * BRO_TREE_U16(BRO_ALPHA_CMD), sym);
        else r = bro_decode_sym(d.in, T_cmd + cat[1].btype * BRO_TREE_U16(BRO_ALPHA_CMD), sym);
        if (r == BRO_SYM_HOLE) return BRO_ST_ParseErrorInsertAndCopyLength;
        if (r == BRO_SYM_EOF) return BRO_ST_UnexpectedEOF;
        uint32_t ie = bro_ic_insert[sym], ce = bro_ic_copy[sym];
        uint32_t insert_len = ie & 0xffffu, copy_len = ce & 0xffffu;
        {
            const uint32_t ib = ie >> 16, cb = ce >> 16;                             // insert extra bits come first
            if (ib + cb <= 25u) {
                // both fields from one window read (the common case: short lengths carry few extra bits)
                if (!bro_read_bits(d.in, ib + cb, extra)) return BRO_ST_UnexpectedEOF;
                insert_len += extra & ((1u << ib) - 1u);
                copy_len += extra >> ib;
            } else {
                if (!bro_read_bits(d.in, ib, extra)) return BRO_ST_UnexpectedEOF;
                insert_len += extra;
                if (!bro_read_bits(d.in, cb, extra)) return BRO_ST_UnexpectedEOF;
                copy_len += extra;
            }
        }
        uint32_t mb_out = d.pos - mb_begin;
        if (mlen < mb_out + insert_len) return BRO_ST_ExceededExpectedBytes;       // src/lib.rs:2036-2039
        // literals (src/lib.rs:1286-1365).  The reference decodes all literals of a command before it emits any, so a
        // decode error inside the run wins over a full output slot: keep decoding (without storing) past the end of
        // the slot and report OutputTooSmall only if the whole run decoded.
#if !defined(BRO_SERIAL) && BRO_W == 32u
        // The run in CHUNKS of literals that share a block (the block accounting of src/lib.rs:1296-1302 done per chunk: a
        // literal that finds its block exhausted reads the block switch first).  A chunk whose code does not depend on the
        // context is decoded LANE-PARALLEL as in bro_commands_simple: lane l looks up the code that would start at bit offset
        // l of the next 32 bits, the codes really present are the chain 0 -> len(0) -> ..., walked with one shuffle per
        // literal.  (Round 2 decoded every literal of a meta-block with block types one at a time, ~1,300 cycles each on a
        // single warp: what bounded the corpus batch -- data/metablock_reset has 460,000 such literals.)
        {
            uint32_t k = 0;
            while (k < insert_len) {
                uint32_t n = insert_len - k;
                if (types0) {
                    if (cat[0].blen == 0u) {
                        if ((st = bro_block_switch(d.in, d.arena, cat[0]))) return st;
                        row = bro_lit_row(sc, d.root10, T_lit, cmap_l, ntl, cat[0].btype);
                        if (n > cat[0].blen + 1u) n = cat[0].blen + 1u;
                        cat[0].blen -= n - 1u;
                    } else {
                        if (n > cat[0].blen) n = cat[0].blen;
                        cat[0].blen -= n;
                    }
                }
                const uint32_t pos0 = d.pos;
                if (row != BRO_ROW_CTX) {
                    const uint16_t* T = T_lit + row * BRO_LIT_STRIDE_U16;
                    const BroRoot r_lit = bro_root_ref(sc.root_lit);
                    uint32_t kk = 0;
                    while (kk < n) {
                        bro_refill(d.in);
                        const uint32_t w2 = bro_shfl(d.in.cur, d.in.wi);                  // the word after the window, not consumed
                        const uint32_t lo = bro_funnel_r(d.in.w0, d.in.w1, d.in.bp);     // stream bits [0, 32)
                        const uint32_t hi = bro_funnel_r(d.in.w1, w2, d.in.bp);          // stream bits [32, 64)
                        const uint32_t mine_bits = bro_funnel_r(lo, hi, lane);
                        const uint32_t e = d.root10 ? (uint32_t)d.root10[mine_bits & 0x3ffu] : bro_root_get(r_lit, mine_bits & 0xffu);
                        const uint32_t len_l = e >> 10;                                   // 0: longer than the root (or a hole)
                        uint32_t want = n - kk, bit = 0, mask = 0, cnt = 0;
                        while (cnt < want) {
                            const uint32_t L = bro_shfl(len_l, bit);
                            if (L == 0u || bit + L > 32u) break;
                            mask |= 1u << bit;
                            bit += L;
                            cnt++;
                            if (bit >= 32u) break;
                        }
                        if (cnt != 0u) {
                            const uint32_t at = pos0 + kk + bro_popc(mask & bro_lanemask_lt());
                            if (((mask >> lane) & 1u) && at < d.cap) d.out[at] = (uint8_t)e;
                            // the two bytes in front of the next literal (a later block may be context-modelled)
                            const uint32_t last = 31u - (uint32_t)__clz(mask), rest = mask & ~(1u << last);
                            const uint32_t b1 = bro_shfl(e, last) & 0xffu, b2 = bro_shfl(e, rest ? 31u - (uint32_t)__clz(rest) : 0u) & 0xffu;
                            d.p2 = rest ? b2 : d.p1; d.p1 = b1;
                            kk += cnt;
                            bro_consume(d.in, bit);
                        } else {
                            uint32_t lit;
                            r = bro_decode_sym_onchip(d.in, r_lit, T, lit);
                            if (r != BRO_SYM_OK) return r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF;
                            if (lane == 0 && pos0 + kk < d.cap) d.out[pos0 + kk] = (uint8_t)lit;
                            d.p2 = d.p1; d.p1 = lit;
                            kk++;
                        }
                    }
                    if ((int32_t)d.in.avail < 0) return BRO_ST_UnexpectedEOF;             // checkpoint (the rounds consume unchecked)
                    d.pos += n;
                } else {
                    // a code per literal chosen by the context; lane kk % W keeps literal kk, W literals leave with one coalesced
                    // store (nothing is stored behind the end of the slot)
                    uint32_t mine = 0;
                    const uint32_t bt = cat[0].btype, mode = modes[bt];
                    for (uint32_t kk = 0; kk < n; kk++) {
                        uint32_t cid;
                        if (mode == 0u) cid = d.p1 & 0x3fu;
                        else if (mode == 1u) cid = d.p1 >> 2;
                        else if (mode == 2u) cid = (uint32_t)bro_lut0[d.p1] | bro_lut1[d.p2];
                        else cid = ((uint32_t)bro_lut2[d.p1] << 3) | bro_lut2[d.p2];
                        const uint32_t t = cmap_l[bt * 64u + cid];
                        const uint16_t* T = T_lit + t * BRO_LIT_STRIDE_U16;
                        uint32_t lit;
                        if (hot.rb_lit) r = bro_decode_sym_hot(d.in, sc.hot_lit + (t << hot.rb_lit), hot.rb_lit, T, lit);
                        else r = bro_decode_sym(d.in, T, lit);
                        if (r == BRO_SYM_HOLE) return BRO_ST_ParseErrorInsertLiterals;
                        if (r == BRO_SYM_EOF) return BRO_ST_UnexpectedEOF;
                        if (lane == (kk & (BRO_W - 1u))) mine = lit;
                        d.pos += 1;
                        d.p2 = d.p1; d.p1 = lit;
                        if (((kk + 1u) & (BRO_W - 1u)) == 0u) {
                            const uint32_t at = pos0 + kk + 1u - BRO_W + lane;
                            if (at < d.cap) d.out[at] = (uint8_t)mine;
                        }
                    }
                    const uint32_t tail = n & (BRO_W - 1u), at = pos0 + n - tail + lane;
                    if (lane < tail && at < d.cap) d.out[at] = (uint8_t)mine;
                }
                k += n;
            }
        }
#else
        if (lit_simple && insert_len <= d.cap - d.pos) {
            // fast path: one literal code, no block switches, the run fits the slot.  Lane k%W keeps literal k and the
            // group stores W literals with one coalesced store.
            uint8_t* o = d.out + d.pos;
            uint32_t mine = 0, k = 0, lit = 0;
            while (k < insert_len) {
                r = bro_decode_sym2(d.in, sc.root_lit, T_lit, lit);
                if (r != BRO_SYM_OK) return r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF;
                if (lane == (k & (BRO_W - 1u))) mine = lit;
                d.p2 = d.p1; d.p1 = lit;
                k++;
                if ((k & (BRO_W - 1u)) == 0u) o[k - BRO_W + lane] = (uint8_t)mine;
            }
            uint32_t tail = k & (BRO_W - 1u);
            if (lane < tail) o[k - tail + lane] = (uint8_t)mine;
            d.pos += insert_len;
        } else {
            // the general run: block switches and a code per literal chosen by the context; lane k % W keeps literal k, W
            // literals leave with one coalesced store (nothing is stored behind the end of the slot)
            const uint32_t pos0 = d.pos;
            uint32_t mine = 0;
            for (uint32_t k = 0; k < insert_len; k++) {
                if (types0 && (st = bro_step_block(d, cat[0]))) return st;
                uint32_t t = 0;
                if (ntl >= 2u) {
                    uint32_t bt = cat[0].btype, mode = modes[bt], cid;
                    if (mode == 0u) cid = d.p1 & 0x3fu;
                    else if (mode == 1u) cid = d.p1 >> 2;
                    else if (mode == 2u) cid = (uint32_t)bro_lut0[d.p1] | bro_lut1[d.p2];
                    else cid = ((uint32_t)bro_lut2[d.p1] << 3) | bro_lut2[d.p2];
                    t = cmap_l[bt * 64u + cid];
                }
                const uint16_t* T = T_lit + t * BRO_LIT_STRIDE_U16;
                uint32_t lit;
                if (one_lit) r = bro_decode_sym2(d.in, sc.root_lit, T, lit);
                else if (hot.rb_lit) r = bro_decode_sym_hot(d.in, sc.hot_lit + (t << hot.rb_lit), hot.rb_lit, T, lit);
                else r = bro_decode_sym(d.in, T, lit);
                if (r == BRO_SYM_HOLE) return BRO_ST_ParseErrorInsertLiterals;
                if (r == BRO_SYM_EOF) return BRO_ST_UnexpectedEOF;
                if (lane == (k & (BRO_W - 1u))) mine = lit;
                d.pos += 1;
                d.p2 = d.p1; d.p1 = lit;
                if (((k + 1u) & (BRO_W - 1u)) == 0u) {
                    const uint32_t at = pos0 + k + 1u - BRO_W + lane;
                    if (at < d.cap) d.out[at] = (uint8_t)mine;
                }
            }
            const uint32_t tail = insert_len & (BRO_W - 1u), at = pos0 + insert_len - tail + lane;
            if (lane < tail && at < d.cap) d.out[at] = (uint8_t)mine;
        }
#endif
        if (d.pos > d.cap) { d.pos = d.cap; return BRO_ST_OutputTooSmall; }
        if (d.pos - mb_begin == mlen) return 0;                                      // src/lib.rs:2069-2070
        // distance code (src/lib.rs:1367-1410)
        uint32_t dcode = 0;
        if (sym >= 128u) {
            if (types2 && (st = bro_step_block(d, cat[2]))) return st;
            uint32_t t = 0;
            if (ntd >= 2u) {
                uint32_t cid = copy_len <= 4u ? copy_len - 2u : 3u;
                t = cmap_d[cat[2].btype * 4u + cid];
            }
            const uint16_t* T = T_dist + t * dist_stride;
            if (one_dist) r = bro_decode_sym2(d.in, sc.root_dist, T_dist, dcode);
            else if (hot.rb_dist) r = bro_decode_sym_hot(d.in, sc.root_dist + (t << hot.rb_dist), hot.rb_dist, T, dcode);
            else r = bro_decode_sym(d.in, T, dcode);
            if (r == BRO_SYM_HOLE) return BRO_ST_ParseErrorDistanceCode;
            if (r == BRO_SYM_EOF) return BRO_ST_UnexpectedEOF;
        }
        // distance (src/lib.rs:1412-1481)
        uint32_t distance, max_allowed;
        if ((st = bro_resolve_distance(d, dcode, npostfix, ndirect, distance, max_allowed))) return st;
        // ---- phase two: materialise the copy ----
        if ((st = bro_emit_copy(d, sc, mlen, mb_begin, distance, max_allowed, copy_len, true))) return st;
        if (d.pos - mb_begin == mlen) return 0;                                      // src/lib.rs:2128-2130
    }
}

// On-chip copies of the root tables of a meta-block that has exactly one code of a kind.
BRO_FN void bro_stage_roots(BroDec& d, BroScratch& sc, const BroMbInfo& mb) {
    const unsigned lane = bro_lane();
    const uint16_t* T_lit = d.arena + mb.o_lit;
    const uint16_t* T_cmd = d.arena + mb.o_cmd;
    const uint16_t* T_dist = d.arena + mb.o_dist;
    const bool one_lit = mb.ntl == 1u, one_cmd = mb.cat[1].nbl == 1u, one_dist = mb.ntd == 1u;
    for (uint32_t r = lane; r < 256u; r += BRO_W) {
        if (one_lit) sc.root_lit[r] = T_lit[r];
        if (one_cmd) sc.root_cmd[r] = T_cmd[r];
        if (one_dist) sc.root_dist[r] = T_dist[r];
    }
    bro_syncwarp();
}

// root bits such that `count` narrow roots fit `budget` entries (8 .. `min_bits`), or 0
BRO_FN uint32_t bro_hot_bits(uint32_t count, uint32_t budget, uint32_t min_bits) {
    for (uint32_t rb = 8u; rb >= min_bits; rb--) if ((count << rb) <= budget) return rb;
    return 0u;
}

// narrow roots of `count` tables (stride apart in the arena) into `dst`
BRO_FN void bro_stage_narrow(uint16_t* dst, const uint16_t* T0, uint32_t stride, uint32_t count, uint32_t rb) {
    for (uint32_t i = bro_lane(); i < (count << rb); i += BRO_W) {
        const uint32_t e = T0[(i >> rb) * stride + (i & ((1u << rb) - 1u))], l = e >> 10;
        dst[i] = (uint16_t)((l >= 1u && l <= rb) ? e : 1u);
    }
}

// The on-chip tables of a meta-block with several codes of a kind (the general loop), as far as they fit.
BRO_FN BroHot bro_stage_hot(BroDec& d, BroScratch& sc, const BroMbInfo& mb) {
    const unsigned lane = bro_lane();
    BroHot hot;
    const uint32_t nl = mb.cat[0].nbl, ni = mb.cat[1].nbl, nd = mb.cat[2].nbl;
    hot.rb_lit = mb.ntl >= 2u ? bro_hot_bits(mb.ntl, BRO_HOT_LIT_U16, 5u) : 0u;
    hot.rb_cmd = ni >= 2u ? bro_hot_bits(ni, 256u, 5u) : 0u;
    hot.rb_dist = mb.ntd >= 2u ? bro_hot_bits(mb.ntd, 256u, 4u) : 0u;
    hot.cmap_l = mb.ntl >= 2u && 64u * nl <= BRO_HOT_CMAP_L;
    hot.cmap_d = mb.ntd >= 2u && 4u * nd <= BRO_HOT_CMAP_D;
    hot.modes = nl <= BRO_HOT_MODES;
    if (hot.rb_lit) bro_stage_narrow(sc.hot_lit, d.arena + mb.o_lit, BRO_LIT_STRIDE_U16, mb.ntl, hot.rb_lit);
    if (hot.rb_cmd) bro_stage_narrow(sc.root_cmd, d.arena + mb.o_cmd, BRO_TREE_U16(BRO_ALPHA_CMD), ni, hot.rb_cmd);
    if (hot.rb_dist) bro_stage_narrow(sc.root_dist, d.arena + mb.o_dist, mb.dist_stride, mb.ntd, hot.rb_dist);
    if (hot.cmap_l) for (uint32_t i = lane; i < 64u * nl; i += BRO_W) sc.hot_cmap_l[i] = ((const uint8_t*)(d.arena + mb.o_cmap_l))[i];
    if (hot.cmap_d) for (uint32_t i = lane; i < 4u * nd; i += BRO_W) sc.hot_cmap_d[i] = ((const uint8_t*)(d.arena + mb.o_cmap_d))[i];
    if (hot.modes) for (uint32_t i = lane; i < nl; i += BRO_W) sc.hot_modes[i] = ((const uint8_t*)(d.arena + mb.o_modes))[i];
    bro_syncwarp();
    return hot;
}

// one compressed meta-block after MLEN / ISUNCOMPRESSED: src/lib.rs:1745-2141
BRO_FN int bro_decode_compressed_metablock(BroDec& d, uint32_t mlen) {
    BroMbInfo mb;
    BroScratch& sc = BRO_DSC(d);
    int st = bro_metablock_tables(d, mb);
    if (st) return st;
    bro_stage_roots(d, sc, mb);
    if (mb.simple)
        return bro_commands_simple(d, sc, mlen, mb.npostfix, mb.ndirect, d.arena + mb.o_lit, d.arena + mb.o_cmd, d.arena + mb.o_dist);
    const BroHot hot = bro_stage_hot(d, sc, mb);
    return bro_commands_general(d, sc, mlen, mb, hot);
}

#endif

// ------------------------------------------------------------------------------------------------------
// one stream: src/lib.rs:1545-2170.  Returns the status; *out_len = bytes produced.
// ------------------------------------------------------------------------------------------------------
#define BRO_MB_COMPRESSED 0x1000   /* bro_next_metablock: a compressed meta-block follows (is_last, mlen set) */
#define BRO_MB_END 0x1001          /* the stream ended cleanly */

// WBITS (src/lib.rs:89-119, 412-418): 0 -> 16; 1+n -> 17+n; 1000+m -> 8+m (m>=2), 17 (m=0), m=1 reserved (Q8)
BRO_FN int bro_stream_header(BroDec& d) {
    uint32_t b, n;
    if (!bro_read_bits(d.in, 1, b)) return BRO_ST_UnexpectedEOF;
    uint32_t wbits = 16;
    if (b) {
        if (!bro_read_bits(d.in, 3, n)) return BRO_ST_UnexpectedEOF;
        if (n) wbits = 17u + n;
        else {
            if (!bro_read_bits(d.in, 3, n)) return BRO_ST_UnexpectedEOF;
            if (n == 1u) return BRO_ST_UnexpectedEOF;
            wbits = n ? 8u + n : 17u;
        }
    }
    d.window = (1u << wbits) - 16u;
    return 0;
}

// Advance to the next compressed meta-block: consumes meta-block headers, skips metadata blocks and copies stored
// blocks on the way (src/lib.rs:1572-1734), and runs the end-of-stream checks (2155-2167) when the stream ends.
// after_last: the meta-block just decoded had ISLAST set.  Returns BRO_MB_COMPRESSED, BRO_MB_END or an error status.
#if !defined(BRO_PARSE)
// Resume point in front of the next meta-block header (bro_records.h)
BRO_FN void bro_checkpoint(const BroDec& d, BroResume* ck, const uint8_t* in_start, uint32_t flags) {
    const uint64_t bits = bro_bits_position(d.in, in_start);
    if (bro_lane() == 0) {
        ck->in_bits = bits; ck->pos = d.pos; ck->window = d.window;
        ck->dist[0] = d.d0; ck->dist[1] = d.d1; ck->dist[2] = d.d2; ck->dist[3] = d.d3;
        ck->p1 = d.p1; ck->p2 = d.p2; ck->flags = flags; ck->reserved = 0;
    }
}
#endif

// CK (resumable decode only): write a resume point to `ck` in front of every meta-block header.
template <bool CK = false>
BRO_FN int bro_next_metablock(BroDec& d, bool after_last, uint32_t& is_last, uint32_t& mlen, BroResume* ck = 0, const uint8_t* in_start = 0) {
    uint32_t b, n, v;
    bool ended = after_last;
    while (!ended) {
#if !defined(BRO_PARSE)
        if (CK) bro_checkpoint(d, ck, in_start, BRO_RESUME_HEADER);
#endif
        if (!CK) {
            // An empty metadata block that starts on a byte boundary is the single byte 0x06 (ISLAST 0, MNIBBLES 3, reserved bit
            // 0, MSKIPBYTES 0, two zero fill bits; src/lib.rs:1617-1683): the reference's data/empty.compressed.17 / .18 hold
            // 65,537 of them in a row, 12 % of the corpus batch's warp time when they took the general route below.
            // (.18: 65,537 blocks of three bytes -- the same header with MSKIPBYTES 1, followed AT ONCE, i.e. from bit 6 on, by the
            // eight bits of the skip length minus one, then two zero fill bits, then the bytes skipped; up to two skipped bytes
            // stay inside one window read)
            bro_refill(d.in);
            while ((bro_avail(d.in) & 7u) == 0u && bro_avail(d.in) >= 8u) {
                const uint32_t pk = bro_peek(d.in);
                uint32_t take = 0;
                if ((pk & 0xffu) == 0x06u) take = 8u;
                else if ((pk & 0xc03fu) == 0x0016u && bro_avail(d.in) >= 16u) {
                    const uint32_t skip = ((pk >> 6) & 0xffu) + 1u;
                    if (skip <= 2u && bro_avail(d.in) >= 16u + 8u * skip) take = 16u + 8u * skip;
                }
                if (take == 0u) break;
                bro_consume(d.in, take);
                bro_refill(d.in);
            }
        }
        if (!bro_read_bits(d.in, 1, is_last)) return BRO_ST_UnexpectedEOF;
        if (is_last) {
            if (!bro_read_bits(d.in, 1, b)) return BRO_ST_UnexpectedEOF;
            if (b) break;                                                             // ISLASTEMPTY
        }
        if (!bro_read_bits(d.in, 2, n)) return BRO_ST_UnexpectedEOF;              // MNIBBLES
        if (n == 3u) {
            // metadata block (src/lib.rs:1617-1683)
            if (!bro_read_bits(d.in, 1, b)) return BRO_ST_UnexpectedEOF;
            if (b) return BRO_ST_NonZeroReservedBit;
            uint32_t skip_bytes, skip = 0;
            if (!bro_read_bits(d.in, 2, skip_bytes)) return BRO_ST_UnexpectedEOF;
            if (skip_bytes) {
                uint32_t last = 0;
                for (uint32_t i = 0; i < skip_bytes; i++) {
                    if (!bro_read_bits(d.in, 8, last)) return BRO_ST_UnexpectedEOF;
                    skip |= last << (d.quirk_spec ? 8u * i : i);                    // sic: << i (Q1, src/lib.rs:463)
                }
                if (skip_bytes > 1u && last == 0u) return BRO_ST_UnexpectedEOF;    // InvalidMSkipLen is remapped (Q2)
                skip += 1;
            }
            bro_read_byte_tail(d.in, v);
            if (v) return BRO_ST_NonZeroFillBit;
            if (skip_bytes) {
                const uint8_t* a = bro_bits_addr(d.in);
                if ((uint64_t)(d.in.end - a) < (uint64_t)skip) return BRO_ST_UnexpectedEOF;
                bro_bits_seek(d.in, a + skip);
            }
            if (is_last) ended = true;
            continue;
        }
        // MLEN (src/lib.rs:469-483)
        uint32_t nibbles = n + 4u;
        if (!bro_read_bits(d.in, 4u * nibbles, mlen)) return BRO_ST_UnexpectedEOF;
        if (nibbles > 4u && (mlen >> ((nibbles - 1u) * 4u)) == 0u) return BRO_ST_NonZeroTrailerNibble;
        mlen += 1;
        if (!is_last) {
            if (!bro_read_bits(d.in, 1, b)) return BRO_ST_UnexpectedEOF;          // ISUNCOMPRESSED
            if (b) {
                // stored block (src/lib.rs:1701-1734): all MLEN bytes are read before any is emitted
                bro_read_byte_tail(d.in, v);
                if (v) return BRO_ST_NonZeroFillBit;
                const uint8_t* a = bro_bits_addr(d.in);
                if ((uint64_t)(d.in.end - a) < (uint64_t)mlen) return BRO_ST_UnexpectedEOF;
                if (mlen > d.cap - d.pos) return BRO_ST_OutputTooSmall;
#if defined(BRO_PARSE)
                if (d.imm) {
                    if (!d.sizing) for (uint32_t i = 0; i < mlen; i++) d.out[d.pos + i] = a[i];      // rare: a stored block between context-modelled ones
                } else if (!bro_rec_push(d, d.pos, mlen, BRO_REC_STORED, (uint32_t)(a - d.in_base))) return BRO_ST_RecordsFull;
#else
                bro_syncwarp();
                bro_copy_far(d.out + d.pos, a, mlen);
#endif
                d.pos += mlen;
                d.p2 = mlen >= 2u ? a[mlen - 2] : d.p1;
                d.p1 = a[mlen - 1];
                bro_bits_seek(d.in, a + mlen);
                continue;
            }
        }
        return BRO_MB_COMPRESSED;
    }
    // StreamEnd (src/lib.rs:2155-2167)
    bro_read_byte_tail(d.in, v);
    if (v) return BRO_ST_NonZeroTrailerBit;
    bro_refill(d.in);
    if (bro_avail(d.in) > 0u) return BRO_ST_ExpectedEndOfStream;
    return BRO_MB_END;
}

#if !defined(BRO_PARSE)
BRO_FN int bro_decode_stream(BroDec& d) {
    int st = bro_stream_header(d);
    if (st) return st;
    uint32_t is_last = 0, mlen = 0;
    bool after_last = false;
    for (;;) {
        st = bro_next_metablock(d, after_last, is_last, mlen);
        if (st == BRO_MB_END) return BRO_ST_OK;
        if (st != BRO_MB_COMPRESSED) return st;
        st = bro_decode_compressed_metablock(d, mlen);
        if (st) return st;
        after_last = is_last != 0u;
    }
}

// The same, from and to a resume point (bro_records.h): start where `ck` says (all-zero = start of stream), write a
// resume point in front of every meta-block header, and at the clean end of the stream.  The status of a call that ran
// out of input (UnexpectedEOF) or room (OutputTooSmall) inside a meta-block describes the attempt, `ck` the last
// boundary it passed.  in_start / in_end: the input of THIS call (ck->in_bits counts from in_start).
BRO_FN int bro_decode_stream_resume(BroDec& d, BroResume* ck, const uint8_t* in_start, const uint8_t* in_end) {
    const BroResume r = *ck;
    bro_syncwarp();                                     // every lane has read the resume point before lane 0 rewrites it
    d.pos = r.pos; d.p1 = r.p1; d.p2 = r.p2;
    bro_bits_init(d.in, in_start, in_end);
    int st;
    if (r.flags & BRO_RESUME_HEADER) {
        d.window = r.window;
        d.d0 = r.dist[0]; d.d1 = r.dist[1]; d.d2 = r.dist[2]; d.d3 = r.dist[3];
        const uint64_t byte = r.in_bits >> 3;
        if (byte > (uint64_t)(in_end - in_start)) return BRO_ST_UnexpectedEOF;
        bro_bits_seek(d.in, in_start + byte);
        const uint32_t bit = (uint32_t)(r.in_bits & 7u);
        if (bit) {
            if (bro_avail(d.in) < bit) return BRO_ST_UnexpectedEOF;
            bro_consume(d.in, bit);
        }
    } else {
        d.d0 = 4; d.d1 = 11; d.d2 = 15; d.d3 = 16;      // src/lib.rs:407-408
        st = bro_stream_header(d);
        if (st) return st;
    }
    uint32_t is_last = 0, mlen = 0;
    bool after_last = (r.flags & BRO_RESUME_LAST) != 0u;
    for (;;) {
        if (after_last) bro_checkpoint(d, ck, in_start, BRO_RESUME_HEADER | BRO_RESUME_LAST);
        st = bro_next_metablock<true>(d, after_last, is_last, mlen, ck, in_start);
        if (st == BRO_MB_END) {
            bro_checkpoint(d, ck, in_start, BRO_RESUME_HEADER | BRO_RESUME_LAST | BRO_RESUME_ENDED);
            return BRO_ST_OK;
        }
        if (st != BRO_MB_COMPRESSED) return st;
        st = bro_decode_compressed_metablock(d, mlen);
        if (st) return st;
        after_last = is_last != 0u;
    }
}
#endif
