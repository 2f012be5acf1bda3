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
// Copies are never materialised here, so a meta-block whose literal context map really depends on the two previous
// bytes (libbrotli quality >= 10) cannot be decoded by this path: the stream is handed to the fused warp kernel's
// retry pass with BRO_ST_NeedFused (as are streams that outgrow the thread arena or their share of the record arena).
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
#define BRO_K_HEADER 6u    // at a meta-block boundary (or before the stream header): structured code
#define BRO_K_DONE 7u      // no stream
#ifndef BRO_PARSE_LITS_PER_ROUND
#define BRO_PARSE_LITS_PER_ROUND 8
#endif

// What bounds this kernel is not instruction issue but the number of L2 / HBM round trips on a lane's critical path:
// one per look-up in a table of the arena, one per access to a spilled variable (the threads' stacks do not fit L1).
// So everything a round touches is kept close: the state below in registers (the canonical limits and bases of the
// literal code included); 6-bit roots of the current insert&copy and distance tables (their alphabets are skewed: the
// codes that matter are short; a miss is settled by the table's own 8-bit root in HBM, one round trip, and only then by
// the canonical search, four) and the insert/copy length table in shared memory; the literal table -- the one table
// too big to keep on chip for 384 streams per SM -- as 8-bit root plus symbols in canonical order in a compact HBM
// array: one round trip per literal, two for a code longer than 8 bits.
#define BRO_RB_LIT 8u
#ifndef BRO_RB_CMD
#define BRO_RB_CMD 6u
#endif
#ifndef BRO_RB_DIST
#define BRO_RB_DIST 6u
#endif
// Two placements of the literal root, both measured on B200 (profiles/r01_kernel_variants.md):
//   default: in the thread's compact HBM array (an L2 round trip per literal) -- 256 B of shared memory per thread
//            (insert&copy and distance roots), 384 streams per SM: best on the headline high-ratio workload;
//   BRO_PARSE_LIT_SMEM: in shared memory -- 768 B per thread, 256 streams per SM: best on literal-heavy streams.
// d.roots (HBM, L2): the 256 symbols of the literal code in canonical order, one byte each (looked up only for a code
// longer than 8 bits) [, the 8-bit literal root]
#if defined(BRO_PARSE_ALL_SMEM)
//   BRO_PARSE_ALL_SMEM: root and symbols in shared memory -- 1,024 B per thread, one CTA of 192 streams per SM.
#define BRO_ROOTS_U16 8u
#define BRO_LIT_ROOT(d) ((d).roots_cd + BRO_ROOTS_CMD + (1u << BRO_RB_CMD) + (1u << BRO_RB_DIST))
#define BRO_LIT_SORTED(d) ((d).roots_cd + BRO_ROOTS_CMD + (1u << BRO_RB_CMD) + (1u << BRO_RB_DIST) + (1u << BRO_RB_LIT))
#define BRO_ROOTS_CD_U16 ((1u << BRO_RB_CMD) + (1u << BRO_RB_DIST) + (1u << BRO_RB_LIT) + 128u)
#elif defined(BRO_PARSE_LIT_SMEM)
#define BRO_ROOTS_U16 128u
#define BRO_LIT_ROOT(d) ((d).roots_cd + BRO_ROOTS_CMD + (1u << BRO_RB_CMD) + (1u << BRO_RB_DIST))
#define BRO_LIT_SORTED(d) ((d).roots)
#define BRO_ROOTS_CD_U16 ((1u << BRO_RB_CMD) + (1u << BRO_RB_DIST) + (1u << BRO_RB_LIT))
#else
#define BRO_ROOTS_U16 (128u + (1u << BRO_RB_LIT))
#define BRO_LIT_ROOT(d) ((d).roots + 128u)
#define BRO_LIT_SORTED(d) ((d).roots)
#define BRO_ROOTS_CD_U16 ((1u << BRO_RB_CMD) + (1u << BRO_RB_DIST))
#endif
// d.roots_cd (shared memory): 6-bit roots of the insert&copy and the distance table [, the 8-bit literal root]
#define BRO_ROOTS_CMD 0u
#define BRO_ROOTS_DIST (1u << BRO_RB_CMD)

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
    uint32_t multi;           // bit c: category c has >= 2 block types; bit 3: one distance table (its root is on chip)
    uint32_t is_last;         // ISLAST of the current meta-block
    uint32_t started;         // the stream header has been read
    uint32_t npostfix, ndirect, o_dist;   // of the current meta-block (copies of BroMbInfo fields, which lives on the stack)
    // the current literal code beyond its root: limit[9..15] and base[9..15] (two per word), maximum length, single symbol
    uint32_t lit_lim[4], lit_base[4], lit_misc;   // lit_misc = max length | single flag << 8 | single symbol << 16
    int st;                   // final status once kind == BRO_K_DONE
};

BRO_FN void bro_parse_begin(BroParse& ps) {
    ps.kind = BRO_K_HEADER; ps.toff_cmd = 0; ps.toff_lit = 0; ps.ins_rem = 0; ps.copy_len = 0; ps.dcode = 0; ps.need_dist = 0;
    ps.mb_begin = 0; ps.mlen = 0; ps.blen0 = ps.blen1 = ps.blen2 = 0; ps.multi = 0; ps.is_last = 0; ps.started = 0; ps.st = 0;
    ps.npostfix = 0; ps.ndirect = 0; ps.o_dist = 0; ps.lit_misc = 0;
    for (int i = 0; i < 4; i++) { ps.lit_lim[i] = 0; ps.lit_base[i] = 0; }
}

BRO_FN void bro_parse_finish(BroParse& ps, int st) { ps.st = st; ps.kind = BRO_K_DONE; }

// Make T the current literal table: root and canonical-order symbols into the thread's compact array, limits and bases
// of the long codes into registers.
BRO_FN void bro_parse_load_lit(BroParse& ps, uint16_t* lit_sorted, uint16_t* lit_root, const uint16_t* T) {
    bro_narrow_root(lit_root, BRO_RB_LIT, T);
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++) {
        // word k holds lengths 9 + 2k (low half) and 10 + 2k (high half); length 16 does not exist: limit 0xffff
        const uint32_t L0 = 9u + 2u * k, L1 = 10u + 2u * k;
        ps.lit_lim[k] = (uint32_t)T[BRO_T_LIMIT + L0] | ((L1 <= 15u ? (uint32_t)T[BRO_T_LIMIT + L1] : 0xffffu) << 16);
        ps.lit_base[k] = (uint32_t)T[BRO_T_BASE + L0] | ((L1 <= 15u ? (uint32_t)T[BRO_T_BASE + L1] : 0u) << 16);
    }
    ps.lit_misc = (uint32_t)T[BRO_T_MAXDEPTH] | (T[BRO_T_SINGLE] ? 0x100u : 0u) | ((uint32_t)T[BRO_T_SINGLE_SYM] << 16);
    uint8_t* sorted = (uint8_t*)lit_sorted;
#pragma unroll 8
    for (uint32_t i = 0; i < 256u; i++) sorted[i] = (uint8_t)T[BRO_T_SORTED + i];
}

// The part of a literal decode behind a root miss (entry without a length): a code longer than 8 bits -- a search of the
// limits held in registers and ONE look-up of the symbol --, a one-symbol code, or no code at all.  Consumes the bits.
BRO_FN int bro_parse_lit_long(BroBits& s, const BroParse& ps, const uint16_t* lit_sorted, uint32_t peek, uint32_t& sym) {
    const uint32_t avail = bro_avail(s);
    uint32_t len = 0;
    if (ps.lit_misc & 0x100u) sym = ps.lit_misc >> 16;                        // one symbol: zero bits
    else {
        const uint32_t x = bro_brev(peek) >> 17;                               // next 15 bits, first bit read most significant
        // smallest L in 9..15 with x < limit[L] (limits grow with L)
        uint32_t L = 16u, base = 0;
#pragma unroll
        for (int k = 3; k >= 0; k--) {
            const uint32_t lo = ps.lit_lim[k] & 0xffffu, hi = ps.lit_lim[k] >> 16;
            if (k < 3 && x < hi) { L = 10u + 2u * (uint32_t)k; base = ps.lit_base[k] >> 16; }
            if (x < lo) { L = 9u + 2u * (uint32_t)k; base = ps.lit_base[k] & 0xffffu; }
        }
        if (L > 15u) return (avail >= (ps.lit_misc & 0xffu) + 1u) ? BRO_SYM_HOLE : BRO_SYM_EOF;
        len = L;
        sym = ((const uint8_t*)lit_sorted)[((int)(int16_t)base + (int)(x >> (15u - L))) & 255];
    }
    if (len > avail) return BRO_SYM_EOF;
    bro_consume(s, len);
    return BRO_SYM_OK;
}

// One literal (same results as bro_decode_sym on the table; src/huffman/tree/mod.rs:63-92).
BRO_FN int bro_parse_decode_lit(BroBits& s, const BroParse& ps, const uint16_t* lit_sorted, const uint16_t* lit_root, uint32_t& sym) {
    bro_refill(s);
    const uint32_t peek = bro_peek(s);
    const uint32_t e = lit_root[peek & 0xffu];
    const uint32_t len = e >> 10;
    sym = e & 0x3ffu;
    if (len == 0u) return bro_parse_lit_long(s, ps, lit_sorted, peek, sym);
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
        return mb.o_lit + t * BRO_TREE_U16(BRO_ALPHA_LIT);
    }
    uint32_t t = 0;
    if (mb.ntd >= 2u) {
        uint32_t cid = ps.copy_len <= 4u ? ps.copy_len - 2u : 3u;
        t = ((const uint8_t*)(d.arena + mb.o_cmap_d))[mb.cat[2].btype * 4u + cid];
    }
    return mb.o_dist + t * mb.dist_stride;
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
        if (c == 0u) { ps.toff_lit = bro_parse_table(d, ps, mb, 0u); bro_parse_load_lit(ps, BRO_LIT_SORTED(d), BRO_LIT_ROOT(d), d.arena + ps.toff_lit); }
        else if (c == 1u) { ps.toff_cmd = bro_parse_table(d, ps, mb, 1u); bro_narrow_root(d.roots_cd + BRO_ROOTS_CMD, BRO_RB_CMD, d.arena + ps.toff_cmd); }
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
            // (a literal context map that depends on the context has ended the stream with BRO_ST_NeedFused)
            ps.multi = (mb.cat[0].nbl >= 2u ? 1u : 0u) | (mb.cat[1].nbl >= 2u ? 2u : 0u) | (mb.cat[2].nbl >= 2u ? 4u : 0u);
            ps.blen0 = mb.cat[0].blen; ps.blen1 = mb.cat[1].blen; ps.blen2 = mb.cat[2].blen;
            ps.toff_cmd = bro_parse_table(d, ps, mb, 1u);
            ps.toff_lit = bro_parse_table(d, ps, mb, 0u);
            bro_parse_load_lit(ps, BRO_LIT_SORTED(d), BRO_LIT_ROOT(d), d.arena + ps.toff_lit);
            bro_narrow_root(d.roots_cd + BRO_ROOTS_CMD, BRO_RB_CMD, d.arena + ps.toff_cmd);
            if (mb.ntd == 1u) { ps.multi |= 8u; bro_narrow_root(d.roots_cd + BRO_ROOTS_DIST, BRO_RB_DIST, d.arena + mb.o_dist); }
            ps.npostfix = mb.npostfix; ps.ndirect = mb.ndirect; ps.o_dist = mb.o_dist;
            ps.kind = BRO_K_CMD;
            return;
        }
    }
    bro_parse_finish(ps, st == BRO_MB_END ? BRO_ST_OK : st);
}

// One ROUND of the machine: every lane that is inside a meta-block advances by (at most) one command -- its
// insert&copy symbol, up to BRO_PARSE_LITS_PER_ROUND of its literals, its distance code, its copy -- in four steps
// that the lanes of a warp execute together, each lane taking part in the steps its state calls for.  There is no
// early exit inside a step (an error parks the lane in BRO_K_DONE), so the lanes meet again after every step.
BRO_FN void bro_parse_round(BroDec& d, BroParse& ps, BroMbInfo& mb) {
    // ---- step 1: insert&copy command symbol and its extra bits ----
    if (ps.kind == BRO_K_CMD && bro_parse_block_step(d, ps, mb, 1u)) {
        uint32_t sym = 0;
        const int r = bro_decode_sym_r(d.in, d.roots_cd + BRO_ROOTS_CMD, BRO_RB_CMD, d.arena + ps.toff_cmd, sym);
        if (r != BRO_SYM_OK) bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertAndCopyLength : BRO_ST_UnexpectedEOF);
        else {
            const uint32_t ie = d.ic[2u * sym], ce = d.ic[2u * sym + 1u];        // bro_ic_insert / bro_ic_copy, interleaved on chip
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
                if (insert_len != 0u) ps.kind = BRO_K_LIT;
                else bro_parse_after_literals(d, ps);
            }
        }
    }
    // ---- step 2: literals ----
    // Fast loop: when the next `fast` literals of a lane can neither overflow its slot nor run into the end of its
    // input (a literal code has at most 15 bits) and no block switch is due, they need no checks at all.  This loop is
    // the critical path of a literal-heavy stream: every instruction in it is paid in full by a single lane.
    {
        uint32_t fast = 0;
        if (ps.kind == BRO_K_LIT) {
            uint32_t n = ps.ins_rem < BRO_PARSE_LITS_PER_ROUND ? ps.ins_rem : (uint32_t)BRO_PARSE_LITS_PER_ROUND;
            if ((ps.multi & 1u) && ps.blen0 < n) n = ps.blen0;     // literals left in the current block (0: a switch is due)
            if (d.pos <= d.cap && n <= d.cap - d.pos && bro_avail(d.in) >= 16u * n) fast = n;
        }
        const uint16_t* const lit_root = BRO_LIT_ROOT(d);
        uint8_t* op = d.out + d.pos;
        uint32_t done = 0;
#pragma unroll 1
        for (uint32_t u = 0; u < BRO_PARSE_LITS_PER_ROUND; u++) {
            if (!bro_any(u < fast)) break;
            if (u < fast) {
                bro_refill(d.in);
                const uint32_t peek = bro_peek(d.in);
                const uint32_t e = lit_root[peek & 0xffu];
                uint32_t len = e >> 10, sym = e;
                if (len == 0u) {
                    // a code longer than 8 bits, a one-symbol code, or no code at all: the general decoder
                    const int r = bro_parse_lit_long(d.in, ps, BRO_LIT_SORTED(d), peek, sym);
                    if (r != BRO_SYM_OK) { bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF); fast = 0; }
                } else bro_consume(d.in, len);
                if (fast) { if (!d.sizing) op[u] = (uint8_t)sym; done = u + 1u; }
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
            const int r = bro_parse_decode_lit(d.in, ps, BRO_LIT_SORTED(d), BRO_LIT_ROOT(d), sym);
            if (r != BRO_SYM_OK) bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF);
            else {
                // a run that does not fit the slot is still decoded: a decode error inside it wins over OutputTooSmall
                if (d.pos < d.cap && !d.sizing) d.out[d.pos] = (uint8_t)sym;
                d.pos += 1;
                if (--ps.ins_rem == 0u) bro_parse_after_literals(d, ps);
            }
        }
    }
    // ---- step 3: distance code ----
    if (ps.kind == BRO_K_DIST && bro_parse_block_step(d, ps, mb, 2u)) {
        uint32_t sym = 0;
        const int r = (ps.multi & 8u) ? bro_decode_sym_r(d.in, d.roots_cd + BRO_ROOTS_DIST, BRO_RB_DIST, d.arena + ps.o_dist, sym)
                                      : bro_decode_sym(d.in, d.arena + bro_parse_table(d, ps, mb, 2u), sym);
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
                    const int n = bro_dict_word(*d.sc, d.dict, d.quirk_spec, copy_len, index, tid);
                    if (n < 0) st = BRO_ST_PanicUppercaseZero;
                    else if (ps.mlen < mb_out + (uint32_t)n) st = BRO_ST_ExceededExpectedBytes;                  // after the transform (Q10)
                    else if ((uint32_t)n > d.cap - d.pos) st = BRO_ST_OutputTooSmall;
                    else {
                        if (!d.sizing) for (uint32_t i = 0; i < (uint32_t)n; i++) d.out[d.pos + i] = d.sc->word[i];
                        d.pos += (uint32_t)n;
                    }
                }
            }
        }
        if (st) bro_parse_finish(ps, st);
        else ps.kind = (d.pos - ps.mb_begin == ps.mlen) ? BRO_K_HEADER : BRO_K_CMD;                              // src/lib.rs:2128-2130
    }
}
