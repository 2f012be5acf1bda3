// bro_parse.h -- PHASE ONE of the two-phase path: ONE THREAD PER STREAM decodes the entropy-coded part of a stream.
//
// Brotli's entropy decode is serial inside a stream, so a warp that works on one stream wastes 31 of its 32 issue
// slots on it (bro_decode_warp_kernel is issue-bound: profiles/r01_ncu_summary.md).  Here every lane of a warp owns
// a different stream, and the command loop of the reference (src/lib.rs:2003-2141) is restated as a machine that the
// lanes of a warp step through TOGETHER: one round advances every lane by one command in four steps (insert&copy
// symbol, a few literals, distance code, copy), each lane taking part in the steps its own state calls for.  The
// lanes therefore share every instruction of a step however different their streams are, and a lane in the middle
// of a long literal run simply sits out the other steps.
//
// What phase one produces:
//   * literals and static-dictionary words (with their transforms) are written straight into the output slot,
//   * every LZ77 back-reference and every stored meta-block becomes BroRec records {dst, len, distance | source offset},
//     one per piece of at most 32 aligned 16-byte vectors (bro_rec_push); phase two (bro_kernels_copy.cu) executes the
//     records of a stream in order, one warp per stream;
//   * in sizing mode (bro_batch_sizes) nothing at all: only the decoded size of every stream.
// A meta-block whose literal context map really depends on the two previous bytes (libbrotli quality >= 10) needs the
// bytes copies produce while it is decoded.  Such a stream is decoded in IMMEDIATE MODE: its thread executes every copy
// itself as soon as it is decoded (no records: phase two has nothing to do for the stream), so the two bytes in front of
// every literal are there -- its own stores, read back.  The mode is entered at the first such meta-block, provided the
// stream has left nothing to phase two so far (otherwise, and in sizing mode, the stream is handed to the fused warp
// kernel's retry pass with BRO_ST_NeedFused, as are streams that outgrow the thread arena or their share of the record
// arena).  Per stream this is slower than the fused kernel's warp; per batch it is what 32 streams per warp buy.
//
// Include with BRO_THREAD_MODE (device) or BRO_HOSTSIM (CPU test-suite) and BRO_PARSE defined.
#pragma once
#if !defined(BRO_PARSE)
#error "bro_parse.h needs BRO_PARSE (and BRO_THREAD_MODE or BRO_HOSTSIM)"
#endif
#include "bro_decoder_core.h"

#define BRO_K_CMD 0u       // next: an insert&copy command symbol            (src/lib.rs:1252-1284)
#define BRO_K_LIT 1u       // next: a literal                                (src/lib.rs:1286-1365)
#define BRO_K_DIST 2u      // next: a distance code                          (src/lib.rs:1367-1410)
#define BRO_K_COPY 3u      // next: distance resolution and the copy itself  (src/lib.rs:1412-1542)
#define BRO_K_LITX 4u      // next: a literal whose code is chosen by the two bytes in front of it (immediate mode)
#define BRO_K_HEADER 6u    // at a meta-block boundary (or before the stream header): structured code
#define BRO_K_DONE 7u      // no stream
#ifndef BRO_PARSE_LITS_PER_ROUND
#define BRO_PARSE_LITS_PER_ROUND 8
#endif
#ifndef BRO_PARSE_LITS_LONG
#define BRO_PARSE_LITS_LONG 64
#endif

// What bounded this kernel in round 1 was not instruction issue but the L2 / HBM round trips on a lane's critical path
// (one per look-up in a table of the arena, one per access to local memory: the threads' stacks did not fit L1), and a
// warp pays for the slowest of its 32 lanes at every step.  So everything a round touches is on chip now, in the
// thread's lane-interleaved block of shared memory (BroTl, bro_decoder_core.h) or in registers:
//   * literals: CANONICAL decode, no root table -- the 15 left-justified limits of the code in registers, the 256
//     symbols in canonical order as bytes in the block.  A literal of ANY length costs ~30 register instructions and
//     ONE conflict-free shared-memory look-up, never a trip to L2 (an 8-bit root would take 512 bytes per stream and
//     still send every longer code off chip: 8 % of the literals of the headline streams);
//   * insert&copy and distance symbols: canonical decode too, from the block alone -- 15 (limit, base) pairs searched
//     by bisection (four look-ups) and the symbols in canonical order, 10 bits each, as many as fit (120 / 24: the
//     shortest codes, i.e. all that matter); only a symbol behind those is fetched from the table in the arena.  Round 1
//     kept 6-bit roots here and settled a miss in the arena: 5 % of the look-ups, but with 32 streams per warp some
//     lane missed in most steps and the whole warp waited for L2 (19 % of the kernel's stall samples);
//   * the insert/copy length table: shared memory, one copy per CTA.
// the block while a meta-block is being decoded (bytes; the header's scratch uses the same bytes, BRO_TL_*)
#define BRO_TLB_LIT 0u                  // uint8[256]: symbols of the current literal code in canonical order
#define BRO_TLB_CMD_TAB 256u            // uint32[16]: the current insert&copy code (bro_parse_load_canon)
#define BRO_TLB_DIST_TAB 320u           // uint32[16]: the distance code (when the meta-block has one)
#define BRO_TLB_CMD_SYMS 384u           // 40 words, three 10-bit symbols each: the first 120 insert&copy symbols in canonical order
#define BRO_TLB_DIST_SYMS 544u          // 8 words: the first 24 distance symbols in canonical order
#define BRO_CMD_ONCHIP 120u
#define BRO_DIST_ONCHIP 24u
static_assert(BRO_TLB_DIST_SYMS + 4u * (BRO_DIST_ONCHIP / 3u) <= BRO_TL_RING, "the decode tables must fit the thread's block");

// The insert&copy length table (src/lookuptable/mod.rs:123, 704 entries) in 48: a symbol's cell (symbol >> 6) gives the
// high parts of its insert and copy length codes, bits 3..5 and 0..2 the low parts; tab[0..24) holds base | extra
// bits << 16 per insert length code, tab[24..48) per copy length code (192 bytes per CTA instead of 5.5 KB).
BRO_FN void bro_ic_compact_entry(uint32_t k, uint32_t& v) {      // k in [0, 48)
    const uint32_t hi = (k % 24u) >> 3, lo = k & 7u;
    v = k < 24u ? bro_ic_insert[(hi == 0u ? 0u : hi == 1u ? 4u : 7u) * 64u + lo * 8u] : bro_ic_copy[(hi == 0u ? 0u : hi == 1u ? 1u : 6u) * 64u + lo];
}
BRO_FN void bro_ic_lookup(const uint32_t* tab, uint32_t sym, uint32_t& ie, uint32_t& ce) {
    const uint32_t cell2 = (sym >> 6) * 2u;
    ie = tab[((0x298500u >> cell2) & 3u) * 8u + ((sym >> 3) & 7u)];
    ce = tab[24u + ((0x262444u >> cell2) & 3u) * 8u + (sym & 7u)];
}

#if defined(BRO_HOSTSIM)
BRO_FN bool bro_any(bool p) { return p; }
#else
BRO_FN bool bro_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
#endif

// per-lane state of the machine (registers)
struct BroParse {
    uint32_t kind;            // BRO_K_*
    uint32_t toff_cmd, toff_lit;   // arena offsets of the insert&copy / literal table of the current block types
    uint32_t ins_rem;         // literals left in the current command
    uint32_t copy_len;
    uint32_t dcode;           // distance code of the current command (0 when it carries none)
    uint32_t need_dist;       // the command carries an explicit distance code (symbol >= 128)
    uint32_t mb_begin, mlen;  // meta-block: output position at its start, MLEN
    uint32_t blen0, blen1, blen2;   // symbols left in the current block per category (valid when the category has >= 2 types)
    uint32_t multi;           // bit c: category c has >= 2 block types; bit 3: one distance table (its root is on chip);
                              // bit 4: literal codes are chosen per context (immediate mode: toff_lit = the context map row of
                              // the current block type), bits 5-6: its context mode, bit 7: the context map is in the block
    uint32_t is_last;         // ISLAST of the current meta-block
    uint32_t started;         // the stream header has been read
    uint32_t npostfix, ndirect, o_dist;   // of the current meta-block (copies of BroMbInfo fields, which lives on the stack)
    // the current literal code: lit_lim[k] = left-justified 15-bit end of all codes of length <= k + 1; lit_dbb[k] =
    // 65536 + base[k + 2] - base[k + 1] (base[L] = canonical index of the first code of length L minus its value), so
    // that ONE sum over the limits the next 15 bits reach yields both the code's length and its base (bro_parse_lit)
    uint32_t lit_lim[15], lit_dbb[15];
    uint32_t lit_misc;        // max length | single flag << 8 | single symbol << 16 | bit 9: no code of 1..5 bits | bit 10: none of 11..15 bits
    uint32_t lit_low;         // sum of lit_dbb[0..5) (what the five shortest lengths contribute when the code has none of them)
    int st;                   // final status once kind == BRO_K_DONE
};

BRO_FN void bro_parse_begin(BroParse& ps) {
    ps.kind = BRO_K_HEADER; ps.toff_cmd = 0; ps.toff_lit = 0; ps.ins_rem = 0; ps.copy_len = 0; ps.dcode = 0; ps.need_dist = 0;
    ps.mb_begin = 0; ps.mlen = 0; ps.blen0 = ps.blen1 = ps.blen2 = 0; ps.multi = 0; ps.is_last = 0; ps.started = 0; ps.st = 0;
    ps.npostfix = 0; ps.ndirect = 0; ps.o_dist = 0; ps.lit_misc = 0; ps.lit_low = 5u * 65536u;
#pragma unroll
    for (int i = 0; i < 15; i++) { ps.lit_lim[i] = 0; ps.lit_dbb[i] = 65536u; }
}

BRO_FN void bro_parse_finish(BroParse& ps, int st) { ps.st = st; ps.kind = BRO_K_DONE; }

#if defined(BRO_HOSTSIM)
// test-suite instrumentation: canonical look-ups [insert&copy | distance][0 all | 1 symbol fetched from the arena | 2 no code]
static uint64_t bro_hostsim_root_stats[2][3];
#define BRO_ROOT_STAT(off, k) (bro_hostsim_root_stats[(off) == BRO_TLB_CMD_TAB ? 0 : 1][k]++)
#else
#define BRO_ROOT_STAT(off, k)
#endif

// Make T the thread's current insert&copy (or distance) code: word l = 1..15 of the table block holds limit[l] (the
// left-justified 15-bit end of all codes of length <= l) | base[l + 1] << 16 (canonical index of the first code of
// length l + 1 minus its value), word 0 what a look-up without a code needs (max length | single flag << 8 | the
// single symbol << 16); the first `cap` symbols of the canonical order go into the symbol block, three to a word.
BRO_FN void bro_parse_load_canon(BroTl t, uint32_t tab, uint32_t syms, uint32_t cap, const uint16_t* T) {
    bro_tl_st32(t, tab, (uint32_t)T[BRO_T_MAXDEPTH] | (T[BRO_T_SINGLE] ? 0x100u : 0u) | ((uint32_t)T[BRO_T_SINGLE_SYM] << 16));
#pragma unroll
    for (uint32_t l = 1; l <= 15u; l++)
        bro_tl_st32(t, tab + 4u * l, (uint32_t)T[BRO_T_LIMIT + l] | (l < 15u ? (uint32_t)T[BRO_T_BASE + l + 1u] << 16 : 0u));
#pragma unroll 4
    for (uint32_t w = 0; w < cap / 3u; w++) {      // (entries behind the code's last symbol are never looked up)
        const uint16_t* p = T + BRO_T_SORTED + 3u * w;
        bro_tl_st32(t, syms + 4u * w, ((uint32_t)p[0] & 0x3ffu) | (((uint32_t)p[1] & 0x3ffu) << 10) | (((uint32_t)p[2] & 0x3ffu) << 20));
    }
}

// One symbol of such a code (same results as bro_decode_sym on the table; src/huffman/tree/mod.rs:63-92): the code's
// length is 1 + the number of limits the next 15 bits reach -- found by bisection, and the last limit reached brings the
// base of the length along -- and its symbol sits at base + (bits >> (15 - length)) of the canonical order.
BRO_FN int bro_decode_sym_canon(BroBits& s, BroTl t, uint32_t tab, uint32_t syms, uint32_t cap, const uint16_t* T, uint32_t& sym) {
    bro_refill(s);
    const uint32_t x = bro_brev(bro_peek(s)) >> 17;
    uint32_t c = 0, base = 0;                                // base[1] = 0: the first code of length 1 is code 0 at index 0
#pragma unroll
    for (uint32_t k = 8u; k; k >>= 1) {
        const uint32_t e = bro_tl_ld32(t, tab + 4u * (c + k));
        if (x >= (e & 0xffffu)) { c += k; base = e >> 16; }
    }
    BRO_ROOT_STAT(tab, 0);
    if (c >= 15u) {
        // no code starts with these bits: a one-symbol code (zero bits), or a hole / the end of the input
        const uint32_t m = bro_tl_ld32(t, tab);
        BRO_ROOT_STAT(tab, 2);
        if (m & 0x100u) { sym = m >> 16; return BRO_SYM_OK; }
        return (bro_avail(s) >= (m & 0xffu) + 1u) ? BRO_SYM_HOLE : BRO_SYM_EOF;
    }
    const uint32_t len = c + 1u;
    if (len > bro_avail(s)) return BRO_SYM_EOF;
    bro_consume(s, len);
    const uint32_t idx = (uint32_t)((int)(int16_t)base + (int)(x >> (14u - c))) & 1023u;
    if (idx < cap) {
        const uint32_t q = (idx * 0xaaabu) >> 17;             // idx / 3
        sym = (bro_tl_ld32(t, syms + 4u * q) >> (10u * (idx - 3u * q))) & 0x3ffu;
    } else { sym = T[BRO_T_SORTED + idx]; BRO_ROOT_STAT(tab, 1); }
    return BRO_SYM_OK;
}

// Make T the current literal table: symbols in canonical order into the thread's block, limits and base differences
// into registers.
BRO_FN void bro_parse_load_lit(BroParse& ps, BroTl t, const uint16_t* T) {
    int prev = (int)(int16_t)T[BRO_T_BASE + 1u];             // = 0: the first code of length 1 is code 0 at index 0
#pragma unroll
    for (uint32_t k = 0; k < 15u; k++) {
        ps.lit_lim[k] = T[BRO_T_LIMIT + 1u + k];
        const int next = k < 14u ? (int)(int16_t)T[BRO_T_BASE + 2u + k] : prev;
        ps.lit_dbb[k] = (uint32_t)(65536 + next - prev);
        prev = next;
    }
    ps.lit_misc = (uint32_t)T[BRO_T_MAXDEPTH] | (T[BRO_T_SINGLE] ? 0x100u : 0u) | ((uint32_t)T[BRO_T_SINGLE_SYM] << 16);
    // Codes of real data use a band of lengths (the headline streams: 7..10 bits).  A length without codes below the band has
    // limit 0 (always reached), one above it the limit 32768 (never reached): when that holds for a whole group of five
    // lengths in every lane of the warp, the group's compares are skipped (bro_parse_lit)
    if ((ps.lit_lim[0] | ps.lit_lim[1] | ps.lit_lim[2] | ps.lit_lim[3] | ps.lit_lim[4]) == 0u) ps.lit_misc |= 0x200u;
    if (ps.lit_lim[9] == 32768u) ps.lit_misc |= 0x400u;      // limits grow with the length: all of 11..15 are 32768 too
    ps.lit_low = ps.lit_dbb[0] + ps.lit_dbb[1] + ps.lit_dbb[2] + ps.lit_dbb[3] + ps.lit_dbb[4];
    // sorted[] starts 8 bytes into a 16-byte granule of the table record: 64 loads of four symbols, eight in flight
#pragma unroll 8
    for (uint32_t i = 0; i < 64u; i++) {
#if defined(__CUDACC__)
        const uint2 v = ((const uint2*)(T + BRO_T_SORTED))[i];
        const uint32_t vx = v.x, vy = v.y;
#else
        const uint32_t vx = ((const uint32_t*)(T + BRO_T_SORTED))[2u * i], vy = ((const uint32_t*)(T + BRO_T_SORTED))[2u * i + 1u];
#endif
        bro_tl_st32(t, BRO_TLB_LIT + 4u * i, (vx & 0xffu) | ((vx >> 8) & 0xff00u) | ((vy & 0xffu) << 16) | ((vy >> 16) << 24));
    }
}

// One literal of the current code from the next bits `peek` (same results as bro_decode_sym on the table;
// src/huffman/tree/mod.rs:63-92).  Canonical decode: the code's length is 1 + the number of limits the next 15 bits
// (first bit most significant) reach, its symbol sits at base[length] + (bits >> (15 - length)) of the canonical order.
// -> BRO_SYM_*; len = bits to consume (not consumed here: the caller knows whether `avail` must be checked).
// Split in two so that a caller can find the codes of several literals before it fetches their symbols (the look-ups then
// overlap): bro_parse_lit_find -> `ref` = position in the canonical order (or 0x80000000 | the symbol of a one-symbol code),
// bro_parse_lit_fetch -> the symbol.
BRO_FN int bro_parse_lit_find(const BroParse& ps, uint32_t peek, uint32_t avail, uint32_t& ref, uint32_t& len,
                              bool skip_low = false, bool skip_high = false) {      // warp-uniform: see bro_parse_load_lit
    const uint32_t x = bro_brev(peek) >> 17;
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0;                   // three partial sums: a shorter dependency chain
    if (skip_low) acc0 = ps.lit_low;
    else {
#pragma unroll
        for (uint32_t k = 0; k < 5u; k++) if (x >= ps.lit_lim[k]) acc0 += ps.lit_dbb[k];
    }
#pragma unroll
    for (uint32_t k = 5; k < 10u; k++) if (x >= ps.lit_lim[k]) acc1 += ps.lit_dbb[k];
    if (!skip_high) {
#pragma unroll
        for (uint32_t k = 10; k < 15u; k++) if (x >= ps.lit_lim[k]) acc2 += ps.lit_dbb[k];
    }
    const uint32_t acc = acc0 + acc1 + acc2;
    const uint32_t count = (acc + 32768u) >> 16;             // limits reached (|base| < 32768)
    if (count >= 15u) {
        // no code starts with these bits: a one-symbol code (zero bits), or a hole / the end of the input
        len = 0;
        if (ps.lit_misc & 0x100u) { ref = 0x80000000u | (ps.lit_misc >> 16); return BRO_SYM_OK; }
        return (avail >= (ps.lit_misc & 0xffu) + 1u) ? BRO_SYM_HOLE : BRO_SYM_EOF;
    }
    len = count + 1u;
    const int base = (int)(acc - (count << 16));
    ref = (uint32_t)(base + (int)(x >> (14u - count))) & 255u;
    return BRO_SYM_OK;
}
BRO_FN uint32_t bro_parse_lit_fetch(BroTl t, uint32_t ref) {
    return (ref & 0x80000000u) ? (ref & 0xffu) : bro_tl_ld8(t, BRO_TLB_LIT + ref);
}
BRO_FN int bro_parse_lit(const BroParse& ps, BroTl t, uint32_t peek, uint32_t avail, uint32_t& sym, uint32_t& len) {
    uint32_t ref = 0;
    const int r = bro_parse_lit_find(ps, peek, avail, ref, len);
    if (r == BRO_SYM_OK) sym = bro_parse_lit_fetch(t, ref);
    return r;
}

// The same with the end-of-input check, consuming the bits.
BRO_FN int bro_parse_decode_lit(BroBits& s, const BroParse& ps, BroTl t, uint32_t& sym) {
    bro_refill(s);
    uint32_t len = 0;
    const int r = bro_parse_lit(ps, t, bro_peek(s), bro_avail(s), sym, len);
    if (r != BRO_SYM_OK) return r;
    if (len > bro_avail(s)) return BRO_SYM_EOF;
    bro_consume(s, len);
    return BRO_SYM_OK;
}

// Table of a symbol of category c (0 literal, 1 insert&copy, 2 distance) under the current block types.
BRO_FN uint32_t bro_parse_table(const BroDec& d, const BroParse& ps, const BroMbInfo& mb, uint32_t c) {
    if (c == 1u) return mb.o_cmd + ((ps.multi & 2u) ? mb.cat[1].btype * BRO_TREE_U16(BRO_ALPHA_CMD) : 0u);
    if (c == 0u) {
        // the context map is constant over the 64 contexts of every block type (checked at the header)
        uint32_t t = mb.ntl >= 2u ? ((const uint8_t*)(d.arena + mb.o_cmap_l))[mb.cat[0].btype * 64u] : 0u;
        return mb.o_lit + t * BRO_LIT_STRIDE_U16;
    }
    uint32_t t = 0;
    if (mb.ntd >= 2u) {
        uint32_t cid = ps.copy_len <= 4u ? ps.copy_len - 2u : 3u;
        t = ((const uint8_t*)(d.arena + mb.o_cmap_d))[mb.cat[2].btype * 4u + cid];
    }
    return mb.o_dist + t * mb.dist_stride;
}

// L2 residency hints for what immediate mode touches (evict_last for the first lines of the literal tables, evict_first for
// the output and the copy sources): measured on B200, 71.3 ms against 67.0 ms without them on c6_text_q11_w16 -- off.
#ifndef BRO_PARSE_L2_HINTS
#define BRO_PARSE_L2_HINTS 0
#endif
#if defined(__CUDACC__) && BRO_PARSE_L2_HINTS
BRO_FN uint32_t bro_ld16_keep(const uint16_t* p) {
    uint64_t pol; uint16_t v;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return v;
}
BRO_FN uint32_t bro_ld8_stream(const uint8_t* p) {
    uint64_t pol; uint32_t v;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
BRO_FN void bro_st8_stream(uint8_t* p, uint32_t v) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}
#else
BRO_FN uint32_t bro_ld16_keep(const uint16_t* p) { return *p; }
BRO_FN uint32_t bro_ld8_stream(const uint8_t* p) { return *p; }
BRO_FN void bro_st8_stream(uint8_t* p, uint32_t v) { *p = (uint8_t)v; }
#endif

// One symbol of a table in the arena whose root was built (literal codes chosen per context).  Same results as
// bro_decode_sym; the look-up starts in the first 64 entries of the root: an entry there is the answer for every code of at
// most 6 bits (a code of length L fills all entries whose low L bits are its bits), so the lines a stream keeps hot are 128
// bytes per table instead of 512 -- the tables of all resident streams then stay in L2.
BRO_FN int bro_parse_decode_sym_arena(BroBits& s, const uint16_t* T, uint32_t& sym) {
    bro_refill(s);
    const uint32_t peek = bro_peek(s);
    uint32_t e = bro_ld16_keep(T + (peek & 63u));
    uint32_t len = e >> 10;
    if (len - 1u >= 6u) { e = T[peek & (BRO_ROOT_SIZE - 1u)]; len = e >> 10; }      // 7 / 8 bits, or longer than the root
    if (len != 0u) {
        if (len > bro_avail(s)) return BRO_SYM_EOF;
        bro_consume(s, len);
        sym = e & 0x3ffu;
        return BRO_SYM_OK;
    }
    const uint32_t r = bro_sym_slow(T, peek, e, bro_avail(s));
    bro_consume(s, (r >> 16) & 0xffu);
    sym = r & 0xffffu;
    return (int)(r >> 24);
}

#define BRO_PM_LCTX 0x10u
#define BRO_PM_CMAP_ONCHIP 0x80u
// Literal codes chosen per context (src/lib.rs:1286-1365): the context map row and the context mode of the current literal
// block type.  The map of up to four block types lives in the thread's block (where the literal symbols of a meta-block
// without context modelling are), a larger one is read from the arena.
BRO_FN void bro_parse_ctx_row(const BroDec& d, BroParse& ps, const BroMbInfo& mb) {
    const uint32_t bt = mb.cat[0].btype;
    const uint32_t mode = ((const uint8_t*)(d.arena + mb.o_modes))[bt] & 3u;
    ps.multi = (ps.multi & ~0x60u) | (mode << 5);
    ps.toff_lit = ((ps.multi & BRO_PM_CMAP_ONCHIP) ? BRO_TLB_LIT : 2u * mb.o_cmap_l) + 64u * bt;
}

// Immediate mode: an LZ77 back-reference executed by the stream's own thread (src/lib.rs:1483-1505).  Up to 8 bytes per trip:
// ALL loads of a trip are issued before its first store (the compiler cannot know that the stores do not feed the loads,
// and a loop that alternates them pays one memory round trip per byte); a source closer than 8 bytes is a period kept in a
// register.  p1 / p2 leave as the last two bytes written (len >= 2: both come from this copy).
BRO_FN void bro_parse_copy_now(uint8_t* o, uint32_t distance, uint32_t len, uint32_t& p1, uint32_t& p2) {
    const uint8_t* s = o - distance;
    uint32_t last = p1, prev = p2;
    if (distance >= 8u) {
#pragma unroll 1
        for (uint32_t k = 0; k < len; k += 8u) {
            const uint32_t m = len - k;                      // bytes of this trip: min(m, 8)
            uint32_t b[8];
#pragma unroll
            for (uint32_t i = 0; i < 8u; i++) b[i] = i < m ? bro_ld8_stream(s + k + i) : 0u;
#pragma unroll
            for (uint32_t i = 0; i < 8u; i++) if (i < m) { bro_st8_stream(o + k + i, b[i]); prev = last; last = b[i]; }
        }
    } else {
        uint64_t pat = 0;
        {
            uint32_t b[7];
#pragma unroll
            for (uint32_t i = 0; i < 7u; i++) b[i] = i < distance ? bro_ld8_stream(s + i) : 0u;
#pragma unroll
            for (uint32_t i = 0; i < 7u; i++) pat |= (uint64_t)b[i] << (8u * i);
        }
        uint32_t j = 0;
#pragma unroll 1
        for (uint32_t k = 0; k < len; k++) {
            const uint32_t b = (uint32_t)(pat >> (8u * j)) & 0xffu;
            bro_st8_stream(o + k, b);
            prev = last; last = b;
            if (++j == distance) j = 0;
        }
    }
    p1 = last; p2 = prev;
}

// Count the next symbol of category c against its current block; when the block is exhausted, read the block switch
// (src/lib.rs:1182-1250).  Rare (blocks are hundreds of symbols long), so the switch itself is an out-of-line call.
// Returns false when the stream ended with an error.
BRO_FN bool bro_parse_block_step(BroDec& d, BroParse& ps, BroMbInfo& mb, uint32_t c) {
    if (!((ps.multi >> c) & 1u)) return true;
    uint32_t bl = c == 0u ? ps.blen0 : c == 1u ? ps.blen1 : ps.blen2;
    if (bl == 0u) {
        BroBits t = d.in;
        BroBlockCat tc = mb.cat[c];
        const int st = bro_block_switch_cold(t, d.arena, tc);
        d.in = t;
        mb.cat[c] = tc;
        if (st) { bro_parse_finish(ps, st); return false; }
        bl = tc.blen + 1u;
        if (c == 0u) {
            if (ps.multi & BRO_PM_LCTX) bro_parse_ctx_row(d, ps, mb);
            else { ps.toff_lit = bro_parse_table(d, ps, mb, 0u); bro_parse_load_lit(ps, d.scv.t, d.arena + ps.toff_lit); }
        }
        else if (c == 1u) { ps.toff_cmd = bro_parse_table(d, ps, mb, 1u); bro_parse_load_canon(d.scv.t, BRO_TLB_CMD_TAB, BRO_TLB_CMD_SYMS, BRO_CMD_ONCHIP, d.arena + ps.toff_cmd); }
    }
    bl -= 1u;
    if (c == 0u) ps.blen0 = bl; else if (c == 1u) ps.blen1 = bl; else ps.blen2 = bl;
    return true;
}

// After the literals of a command (src/lib.rs:2060-2101).
BRO_FN void bro_parse_after_literals(BroDec& d, BroParse& ps) {
    if (d.pos > d.cap) { d.pos = d.cap; bro_parse_finish(ps, BRO_ST_OutputTooSmall); return; }
    if (d.pos - ps.mb_begin == ps.mlen) { ps.kind = BRO_K_HEADER; return; }                    // src/lib.rs:2069-2070
    ps.kind = ps.need_dist ? BRO_K_DIST : BRO_K_COPY;
}

// Header work at a meta-block boundary: (stream header,) meta-block headers up to the next compressed meta-block,
// its block-type codes, context maps and prefix code tables.  Structured code, the same as the fused kernel runs.
BRO_FN void bro_parse_header(BroDec& d, BroParse& ps, BroMbInfo& mb) {
    int st = 0;
    if (!ps.started) {
        ps.started = 1;
        st = bro_stream_header(d);
    }
    uint32_t is_last = ps.is_last, mlen = 0;
    if (!st) st = bro_next_metablock(d, ps.is_last != 0u, is_last, mlen);
    if (st == BRO_MB_COMPRESSED) {
        ps.is_last = is_last; ps.mlen = mlen; ps.mb_begin = d.pos;
        st = bro_metablock_tables(d, mb);
        if (!st) {
            // (a literal context map that depends on the context has switched the stream to immediate mode -- mb.lctx --
            // or ended it with BRO_ST_NeedFused)
            ps.multi = (mb.cat[0].nbl >= 2u ? 1u : 0u) | (mb.cat[1].nbl >= 2u ? 2u : 0u) | (mb.cat[2].nbl >= 2u ? 4u : 0u);
            ps.blen0 = mb.cat[0].blen; ps.blen1 = mb.cat[1].blen; ps.blen2 = mb.cat[2].blen;
            ps.toff_cmd = bro_parse_table(d, ps, mb, 1u);
            if (mb.lctx) {
                // immediate mode (the stream's copies so far were executed by this thread, or there were none): the
                // context map into the block if it fits, the two bytes in front of the meta-block from the output
                ps.multi |= BRO_PM_LCTX;
                if (mb.cat[0].nbl <= 4u) {
                    ps.multi |= BRO_PM_CMAP_ONCHIP;
                    for (uint32_t w = 0; w < 16u * mb.cat[0].nbl; w++)
                        bro_tl_st32(d.scv.t, BRO_TLB_LIT + 4u * w, ((const uint32_t*)(d.arena + mb.o_cmap_l))[w]);
                }
                bro_parse_ctx_row(d, ps, mb);
                d.p1 = d.pos >= 1u ? d.out[d.pos - 1u] : 0u;
                d.p2 = d.pos >= 2u ? d.out[d.pos - 2u] : 0u;
            } else {
                ps.toff_lit = bro_parse_table(d, ps, mb, 0u);
                bro_parse_load_lit(ps, d.scv.t, d.arena + ps.toff_lit);
            }
            bro_parse_load_canon(d.scv.t, BRO_TLB_CMD_TAB, BRO_TLB_CMD_SYMS, BRO_CMD_ONCHIP, d.arena + ps.toff_cmd);
            if (mb.ntd == 1u) { ps.multi |= 8u; bro_parse_load_canon(d.scv.t, BRO_TLB_DIST_TAB, BRO_TLB_DIST_SYMS, BRO_DIST_ONCHIP, d.arena + mb.o_dist); }
            ps.npostfix = mb.npostfix; ps.ndirect = mb.ndirect; ps.o_dist = mb.o_dist;
            ps.kind = BRO_K_CMD;
            return;
        }
    }
    bro_parse_finish(ps, st == BRO_MB_END ? BRO_ST_OK : st);
}

// Static dictionary word + transform (src/lib.rs:1506-1540, src/transformation/mod.rs:84-209) for one thread: the
// geometry first (transformed length, or -1 where the reference panics: uppercase_first on a 0x00 byte), so
// that the caller can run its checks; then the bytes straight into the output slot -- no staging buffer, the UTF-8 walk
// of the uppercase transforms is applied as the bytes pass.  Same results as bro_dict_word.
BRO_FN int bro_parse_dict_geometry(const uint8_t* dict, int quirk_spec, uint32_t copy_len, uint32_t index, uint32_t tid, uint32_t& from, uint32_t& wl) {
    const uint32_t type = bro_xf_type[tid];
    from = 0; wl = copy_len;
    if (type >= 3u && type <= 11u) {            // OmitFirstN: base_word[min(N, len-1)..] (Q3) / spec: [min(N,len)..]
        const uint32_t n = type - 2u;
        from = quirk_spec ? (n < wl ? n : wl) : (n < wl - 1u ? n : wl - 1u);
        wl -= from;
    } else if (type >= 12u) {                   // OmitLastN: base_word[..max(N,len)-N]
        const uint32_t n = type - 11u;
        wl = (wl > n ? wl : n) - n;
    }
    if (type == 1u && !quirk_spec) {
        // what the reference looks at is the byte behind the prefix: the word's first byte, or -- for an empty word,
        // which uppercase_first cannot get -- nothing
        if (dict[bro_dict_offsets[copy_len] + index * copy_len] == 0u) return -1;
    }
    return (int)(bro_xf_prefix_len[tid] + wl + bro_xf_suffix_len[tid]);
}

BRO_FN void bro_parse_dict_emit(uint8_t* o, const uint8_t* dict, uint32_t copy_len, uint32_t index, uint32_t tid, uint32_t from, uint32_t wl) {
    const uint32_t type = bro_xf_type[tid], plen = bro_xf_prefix_len[tid], slen = bro_xf_suffix_len[tid];
    const uint8_t* w = dict + bro_dict_offsets[copy_len] + index * copy_len + from;
    for (uint32_t i = 0; i < plen; i++) o[i] = bro_xf_strings[bro_xf_prefix_off[tid] + i];
    o += plen;
    // uppercase_first (src/transformation/mod.rs:42-82) / uppercase_all (3-40): position `dec` decides -- an ASCII
    // letter is flipped itself, a 2- / 3-byte UTF-8 lead flips bit 5 of the next / bits 0 and 2 of the byte after the
    // next -- and the walk continues behind what it touched (uppercase_first: one decision only)
    uint32_t dec = (type == 1u || type == 2u) ? 0u : 0xffffffffu, tpos = 0xffffffffu, txor = 0;
    for (uint32_t i = 0; i < wl; i++) {
        uint32_t c = w[i];
        if (i == dec) {
            if (c < 192u) { if (c >= 97u && c <= 122u) c ^= 32u; dec = i + 1u; }
            else if (c < 224u) { tpos = i + 1u; txor = 32u; dec = i + 2u; }
            else { tpos = i + 2u; txor = 5u; dec = i + 3u; }
            if (type == 1u) dec = 0xffffffffu;
        } else if (i == tpos) c ^= txor;
        o[i] = (uint8_t)c;
    }
    o += wl;
    for (uint32_t i = 0; i < slen; i++) o[i] = bro_xf_strings[bro_xf_suffix_off[tid] + i];
}

// One ROUND of the machine: every lane that is inside a meta-block advances by (at most) one command -- its
// insert&copy symbol, up to BRO_PARSE_LITS_PER_ROUND of its literals, its distance code, its copy -- in four steps
// that the lanes of a warp execute together, each lane taking part in the steps its state calls for.  There is no
// early exit inside a step (an error parks the lane in BRO_K_DONE), so the lanes meet again after every step.
// IMM = false: no lane of the warp is in immediate mode (warp-uniform, known after the headers): the code of that mode is
// compiled out, so that batches without context modelling run exactly the loop they ran before the mode existed.
template <bool IMM>
BRO_FN void bro_parse_round(BroDec& d, BroParse& ps, BroMbInfo& mb) {
    // ---- step 1: insert&copy command symbol and its extra bits ----
    if (ps.kind == BRO_K_CMD && bro_parse_block_step(d, ps, mb, 1u)) {
        uint32_t sym = 0;
        const int r = bro_decode_sym_canon(d.in, d.scv.t, BRO_TLB_CMD_TAB, BRO_TLB_CMD_SYMS, BRO_CMD_ONCHIP, d.arena + ps.toff_cmd, sym);
        if (r != BRO_SYM_OK) bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertAndCopyLength : BRO_ST_UnexpectedEOF);
        else {
            uint32_t ie, ce;
            bro_ic_lookup(d.ic, sym, ie, ce);
            uint32_t insert_len = ie & 0xffffu, copy_len = ce & 0xffffu, extra = 0, extra2 = 0;
            const uint32_t ib = ie >> 16, cb = ce >> 16;                   // insert extra bits come first
            bool ok;
            if (ib + cb <= 25u) {
                // both fields from one window read (the common case: short lengths carry few extra bits)
                uint32_t v = 0;
                ok = bro_read_bits(d.in, ib + cb, v);
                extra = v & ((1u << ib) - 1u);
                extra2 = v >> ib;
            } else ok = bro_read_bits(d.in, ib, extra) && bro_read_bits(d.in, cb, extra2);
            insert_len += extra;
            copy_len += extra2;
            if (!ok) bro_parse_finish(ps, BRO_ST_UnexpectedEOF);
            else if (ps.mlen < (d.pos - ps.mb_begin) + insert_len) bro_parse_finish(ps, BRO_ST_ExceededExpectedBytes);   // src/lib.rs:2036-2039
            else {
                ps.ins_rem = insert_len; ps.copy_len = copy_len; ps.need_dist = sym >= 128u; ps.dcode = 0;
                if (insert_len != 0u) ps.kind = (IMM && (ps.multi & BRO_PM_LCTX)) ? BRO_K_LITX : BRO_K_LIT;
                else bro_parse_after_literals(d, ps);
            }
        }
    }
    // ---- step 2: literals ----
    // Fast loop: when the next `fast` literals of a lane can neither overflow its slot nor run into the end of its
    // input (a literal code has at most 15 bits) and no block switch is due, they need no checks at all.  This loop is
    // the critical path of a literal-heavy stream: every instruction in it is paid in full by a single lane.
    {
        // a round serves up to BRO_PARSE_LITS_PER_ROUND literals per lane, so that lanes waiting for their distance code or
        // their copy are not held up by a neighbour's long run -- unless no lane waits: then long runs go on (the round's
        // other steps are pure overhead for a stream that is all literals)
        const uint32_t per_round = bro_any(ps.kind == BRO_K_DIST || ps.kind == BRO_K_COPY) ? (uint32_t)BRO_PARSE_LITS_PER_ROUND : (uint32_t)BRO_PARSE_LITS_LONG;
        uint32_t fast = 0;
        if (ps.kind == BRO_K_LIT) {
            uint32_t n = ps.ins_rem < per_round ? ps.ins_rem : per_round;
            if ((ps.multi & 1u) && ps.blen0 < n) n = ps.blen0;     // literals left in the current block (0: a switch is due)
            const uint32_t room = d.pos <= d.cap ? d.cap - d.pos : 0u, safe = bro_avail(d.in) >> 4;
            n = n < room ? n : room;
            fast = n < safe ? n : safe;      // (whatever is left over -- or everything, when this is 0 -- is the general loop's)
        }
        const BroTl tl = d.scv.t;
        uint8_t* op = d.out + d.pos;
        uint32_t done = 0;
        const bool skip_low = !bro_any(fast != 0u && !(ps.lit_misc & 0x200u)), skip_high = !bro_any(fast != 0u && !(ps.lit_misc & 0x400u));
        // two literals per slide of the window: after bro_refill the window holds 64 bits from a bit position < 32, and a
        // literal code has at most 15 bits, so the second literal of a pair still finds its bits (bro_peek_wide)
#pragma unroll 1
        for (uint32_t u = 0; u < per_round; u += 2u) {
            if (!bro_any(u < fast)) break;
            if (u < fast) {
                bro_refill(d.in);
                uint32_t ref0 = 0, ref1 = 0, len = 0;
                int r = bro_parse_lit_find(ps, bro_peek(d.in), bro_avail(d.in), ref0, len, skip_low, skip_high);
                if (r == BRO_SYM_OK) {
                    bro_consume(d.in, len);
                    done = u + 1u;
                    if (u + 1u < fast) {
                        r = bro_parse_lit_find(ps, bro_peek_wide(d.in), bro_avail(d.in), ref1, len, skip_low, skip_high);
                        if (r == BRO_SYM_OK) { bro_consume(d.in, len); done = u + 2u; }
                    }
                    // both symbols are fetched once both codes are known: the two look-ups overlap
                    if (!d.sizing) {
                        const uint32_t sym0 = bro_parse_lit_fetch(tl, ref0), sym1 = done == u + 2u ? bro_parse_lit_fetch(tl, ref1) : 0u;
                        op[u] = (uint8_t)sym0;
                        if (done == u + 2u) op[u + 1u] = (uint8_t)sym1;
                    }
                }
                if (r != BRO_SYM_OK) { bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF); fast = 0; }
            }
        }
        if (done != 0u && ps.kind == BRO_K_LIT) {
            d.pos += done;
            ps.ins_rem -= done;
            if (ps.multi & 1u) ps.blen0 -= done;
            if (ps.ins_rem == 0u) bro_parse_after_literals(d, ps);
        }
    }
    // General loop: the lanes the fast loop could not take (near the end of the slot or of the input, block switches)
#pragma unroll 1
    for (int u = 0; u < BRO_PARSE_LITS_PER_ROUND; u++) {
        const bool slow = ps.kind == BRO_K_LIT && (((ps.multi & 1u) && ps.blen0 == 0u) || d.pos > d.cap || ps.ins_rem > d.cap - d.pos || bro_avail(d.in) < 16u * BRO_PARSE_LITS_PER_ROUND);
        if (!bro_any(slow)) break;
        if (slow && bro_parse_block_step(d, ps, mb, 0u)) {
            uint32_t sym = 0;
            const int r = bro_parse_decode_lit(d.in, ps, d.scv.t, sym);
            if (r != BRO_SYM_OK) bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF);
            else {
                // a run that does not fit the slot is still decoded: a decode error inside it wins over OutputTooSmall
                if (d.pos < d.cap && !d.sizing) d.out[d.pos] = (uint8_t)sym;
                d.pos += 1;
                if (--ps.ins_rem == 0u) bro_parse_after_literals(d, ps);
            }
        }
    }
    // Context loop (immediate mode): the code of every literal is chosen by the two bytes in front of it (src/lib.rs:1286-1365);
    // its table is looked up in the arena (8-bit root + canonical search, as the fused kernel does)
    if (IMM && bro_any(ps.kind == BRO_K_LITX)) {
        const uint32_t per_round = bro_any(ps.kind == BRO_K_DIST || ps.kind == BRO_K_COPY) ? (uint32_t)BRO_PARSE_LITS_PER_ROUND : (uint32_t)BRO_PARSE_LITS_LONG;
#pragma unroll 1
        for (uint32_t u = 0; u < per_round; u++) {
            const bool cx = ps.kind == BRO_K_LITX;
            if (!bro_any(cx)) break;
            if (cx && bro_parse_block_step(d, ps, mb, 0u)) {
                const uint32_t mode = (ps.multi >> 5) & 3u;
                uint32_t cid;
                if (mode == 0u) cid = d.p1 & 0x3fu;
                else if (mode == 1u) cid = d.p1 >> 2;
                else if (mode == 2u) cid = (uint32_t)bro_lut0[d.p1] | bro_lut1[d.p2];
                else cid = ((uint32_t)bro_lut2[d.p1] << 3) | bro_lut2[d.p2];
                const uint32_t t = (ps.multi & BRO_PM_CMAP_ONCHIP) ? bro_tl_ld8(d.scv.t, ps.toff_lit + cid) : ((const uint8_t*)d.arena)[ps.toff_lit + cid];
                uint32_t sym = 0;
                const int r = bro_parse_decode_sym_arena(d.in, d.arena + mb.o_lit + t * BRO_LIT_STRIDE_U16, sym);
                if (r != BRO_SYM_OK) bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF);
                else {
                    // a run that does not fit the slot is still decoded: a decode error inside it wins over OutputTooSmall
                    if (d.pos < d.cap) bro_st8_stream(d.out + d.pos, sym);
                    d.pos += 1;
                    d.p2 = d.p1; d.p1 = sym;
                    if (--ps.ins_rem == 0u) bro_parse_after_literals(d, ps);
                }
            }
        }
    }
    // ---- step 3: distance code ----
    if (ps.kind == BRO_K_DIST && bro_parse_block_step(d, ps, mb, 2u)) {
        uint32_t sym = 0;
        const int r = (ps.multi & 8u) ? bro_decode_sym_canon(d.in, d.scv.t, BRO_TLB_DIST_TAB, BRO_TLB_DIST_SYMS, BRO_DIST_ONCHIP, d.arena + ps.o_dist, sym)
                                      : bro_parse_decode_sym_arena(d.in, d.arena + bro_parse_table(d, ps, mb, 2u), sym);
        if (r != BRO_SYM_OK) bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorDistanceCode : BRO_ST_UnexpectedEOF);
        else { ps.dcode = sym; ps.kind = BRO_K_COPY; }
    }
    // ---- step 4: distance, then the copy: a record for phase two, or a dictionary word emitted here ----
    if (ps.kind == BRO_K_COPY) {
        uint32_t distance = 0, max_allowed = 0;
        int st = bro_resolve_distance(d, ps.dcode, ps.npostfix, ps.ndirect, distance, max_allowed);
        const uint32_t mb_out = d.pos - ps.mb_begin, copy_len = ps.copy_len;
        if (!st) {
            if (distance <= max_allowed) {
                // an LZ77 back-reference: phase two materialises it
                if (ps.mlen < mb_out + copy_len) st = BRO_ST_ExceededExpectedBytes;                              // src/lib.rs:2105-2108
                else if (copy_len > d.cap - d.pos) st = BRO_ST_OutputTooSmall;
                else if (IMM && d.imm) {
                    // immediate mode: the copy itself; it leaves the two bytes the next literal's context is made of (copy_len >= 2)
                    bro_parse_copy_now(d.out + d.pos, distance, copy_len, d.p1, d.p2);
                    d.pos += copy_len;
                }
                else if (!bro_rec_push(d, d.pos, copy_len, BRO_REC_LZ, distance)) st = BRO_ST_RecordsFull;
                else d.pos += copy_len;
            } else if (copy_len < 4u || copy_len > 24u) st = BRO_ST_InvalidLengthInStaticDictionary;
            else {
                // a static dictionary word (src/lib.rs:1506-1540): emitted here, it needs no earlier output
                const uint32_t word_id = distance - max_allowed - 1u;
                const uint32_t bits = bro_dict_size_bits[copy_len];
                const uint32_t index = word_id & ((1u << bits) - 1u), tid = word_id >> bits;
                if (tid > 120u) st = BRO_ST_InvalidTransformId;
                else {
                    uint32_t from = 0, wl = 0;
                    const int n = bro_parse_dict_geometry(d.dict, d.quirk_spec, copy_len, index, tid, from, wl);
                    if (n < 0) st = BRO_ST_PanicUppercaseZero;
                    else if (ps.mlen < mb_out + (uint32_t)n) st = BRO_ST_ExceededExpectedBytes;                  // after the transform (Q10)
                    else if ((uint32_t)n > d.cap - d.pos) st = BRO_ST_OutputTooSmall;
                    else {
                        if (!d.sizing) bro_parse_dict_emit(d.out + d.pos, d.dict, copy_len, index, tid, from, wl);
                        d.pos += (uint32_t)n;
                        if (IMM && (ps.multi & BRO_PM_LCTX)) {
                            d.p1 = d.pos >= 1u ? d.out[d.pos - 1u] : 0u;
                            d.p2 = d.pos >= 2u ? d.out[d.pos - 2u] : 0u;
                        }
                    }
                }
            }
        }
        if (st) bro_parse_finish(ps, st);
        else ps.kind = (d.pos - ps.mb_begin == ps.mlen) ? BRO_K_HEADER : BRO_K_CMD;                              // src/lib.rs:2128-2130
    }
}
