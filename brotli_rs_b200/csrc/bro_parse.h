// bro_parse.h -- PHASE ONE of the two-phase path: ONE THREAD PER STREAM decodes the entropy-coded part of a stream.
//
// Brotli's entropy decode is serial inside a stream, so a warp that works on one stream wastes 31 of its 32 issue
// slots on it (bro_decode_warp_kernel is issue-bound: profiles/r01_ncu_summary.md).  Here every lane of a warp owns
// a different stream, and the command loop of the reference (src/lib.rs:2003-2141) is restated as a FLAT STATE
// MACHINE: one trip of the loop decodes exactly one prefix-code symbol of whatever kind the lane needs next
// (insert&copy command, literal, distance, block type, block count), so the lanes of a warp stay converged on the
// expensive part (window slide, root lookup, canonical search) however different their streams are, and diverge only
// for the few instructions that interpret the symbol.
//
// What phase one produces:
//   * literals and static-dictionary words (with their transforms) are written straight into the output slot,
//   * every LZ77 back-reference and every stored meta-block becomes one BroRec {dst, len, distance | source offset};
//     phase two (bro_kernels_copy.cu) executes the records of a stream in order, one warp per stream.
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

#define BRO_K_CMD 0u       // next symbol: insert&copy command            (src/lib.rs:1252-1284)
#define BRO_K_LIT 1u       // a literal                                   (src/lib.rs:1286-1365)
#define BRO_K_DIST 2u      // a distance code                             (src/lib.rs:1367-1410)
#define BRO_K_COPY0 3u     // no symbol: implicit distance code 0, then the copy
#define BRO_K_BTYPE 4u     // a block type code                           (src/lib.rs:1226-1250)
#define BRO_K_BCOUNT 5u    // a block count code                          (src/lib.rs:957-987)
#define BRO_K_HEADER 6u    // at a meta-block boundary (or before the stream header): structured code
#define BRO_K_DONE 7u      // no stream

// per-lane state of the flat machine (registers)
struct BroParse {
    uint32_t kind;            // BRO_K_*
    uint32_t pending;         // kind to resume after a block switch
    uint32_t cat;             // category of the block switch in progress (0 literals, 1 insert&copy, 2 distances)
    uint32_t toff;            // arena offset of the table the next symbol is decoded with
    uint32_t ins_rem;         // literals left in the current command
    uint32_t copy_len;
    uint32_t need_dist;       // the command carries an explicit distance code (symbol >= 128)
    uint32_t mb_begin, mlen;  // meta-block: output position at its start, MLEN
    uint32_t blen0, blen1, blen2;   // symbols left in the current block per category (valid when the category has >= 2 types)
    uint32_t multi;           // bit c: category c has >= 2 block types
    uint32_t is_last;         // ISLAST of the current meta-block
    uint32_t started;         // the stream header has been read
    int st;                   // final status once kind == BRO_K_DONE
};

BRO_FN void bro_parse_begin(BroParse& ps) {
    ps.kind = BRO_K_HEADER; ps.pending = 0; ps.cat = 0; ps.toff = 0; ps.ins_rem = 0; ps.copy_len = 0; ps.need_dist = 0;
    ps.mb_begin = 0; ps.mlen = 0; ps.blen0 = ps.blen1 = ps.blen2 = 0; ps.multi = 0; ps.is_last = 0; ps.started = 0; ps.st = 0;
}

BRO_FN void bro_parse_finish(BroParse& ps, int st) { ps.st = st; ps.kind = BRO_K_DONE; }

// Table of the next symbol of kind `k` under the current block types.
BRO_FN uint32_t bro_parse_table(const BroDec& d, const BroParse& ps, const BroMbInfo& mb, uint32_t k) {
    if (k == BRO_K_CMD) return mb.o_cmd + ((ps.multi & 2u) ? mb.cat[1].btype * BRO_TREE_U16(BRO_ALPHA_CMD) : 0u);
    if (k == BRO_K_LIT) {
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

// Make `k` (CMD, LIT or DIST, category c) the next symbol: count it against the current block of its category, or
// start a block switch first (src/lib.rs:1182-1197).
BRO_FN void bro_parse_enter(const BroDec& d, BroParse& ps, const BroMbInfo& mb, uint32_t k) {
    const uint32_t c = k == BRO_K_CMD ? 1u : k == BRO_K_LIT ? 0u : 2u;
    if ((ps.multi >> c) & 1u) {
        uint32_t bl = c == 0u ? ps.blen0 : c == 1u ? ps.blen1 : ps.blen2;
        if (bl == 0u) {
            ps.pending = k; ps.cat = c; ps.kind = BRO_K_BTYPE; ps.toff = mb.cat[c].t_type;
            return;
        }
        bl -= 1u;
        if (c == 0u) ps.blen0 = bl; else if (c == 1u) ps.blen1 = bl; else ps.blen2 = bl;
    }
    ps.kind = k;
    ps.toff = bro_parse_table(d, ps, mb, k);
}

// After the literals of a command (src/lib.rs:2060-2101).
BRO_FN void bro_parse_after_literals(BroDec& d, BroParse& ps, const BroMbInfo& mb) {
    if (d.pos > d.cap) { d.pos = d.cap; bro_parse_finish(ps, BRO_ST_OutputTooSmall); return; }
    if (d.pos - ps.mb_begin == ps.mlen) { ps.kind = BRO_K_HEADER; return; }                    // src/lib.rs:2069-2070
    if (ps.need_dist) bro_parse_enter(d, ps, mb, BRO_K_DIST);
    else ps.kind = BRO_K_COPY0;
}

// Header work at a meta-block boundary: (stream header,) meta-block headers up to the next compressed meta-block,
// its block-type codes, context maps and prefix code tables.  Structured code, the same as the fused kernel runs.
BRO_FN void bro_parse_header(BroDec& d, BroParse& ps, BroMbInfo& mb) {
    int st;
    if (!ps.started) {
        ps.started = 1;
        if ((st = bro_stream_header(d))) { bro_parse_finish(ps, st); return; }
    }
    uint32_t is_last = ps.is_last, mlen = 0;
    st = bro_next_metablock(d, ps.is_last != 0u, is_last, mlen);
    if (st == BRO_MB_END) { bro_parse_finish(ps, BRO_ST_OK); return; }
    if (st != BRO_MB_COMPRESSED) { bro_parse_finish(ps, st); return; }
    ps.is_last = is_last; ps.mlen = mlen; ps.mb_begin = d.pos;
    if ((st = bro_metablock_tables(d, mb))) { bro_parse_finish(ps, st); return; }
    // (a literal context map that depends on the context has already ended the stream with BRO_ST_NeedFused)
    ps.multi = (mb.cat[0].nbl >= 2u ? 1u : 0u) | (mb.cat[1].nbl >= 2u ? 2u : 0u) | (mb.cat[2].nbl >= 2u ? 4u : 0u);
    ps.blen0 = mb.cat[0].blen; ps.blen1 = mb.cat[1].blen; ps.blen2 = mb.cat[2].blen;
    bro_parse_enter(d, ps, mb, BRO_K_CMD);
}

// One trip of the flat machine for a lane whose kind is CMD .. BCOUNT.
BRO_FN void bro_parse_step(BroDec& d, BroParse& ps, BroMbInfo& mb) {
    uint32_t sym = 0;
    int r = BRO_SYM_OK;
    const uint32_t kind = ps.kind;
    if (kind != BRO_K_COPY0) r = bro_decode_sym(d.in, d.arena + ps.toff, sym);
    if (kind == BRO_K_LIT) {
        if (r != BRO_SYM_OK) { bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertLiterals : BRO_ST_UnexpectedEOF); return; }
        // a run that does not fit the slot is still decoded: a decode error inside it wins over OutputTooSmall
        if (d.pos < d.cap) d.out[d.pos] = (uint8_t)sym;
        d.pos += 1;
        if (--ps.ins_rem != 0u) {
            if (ps.multi & 1u) bro_parse_enter(d, ps, mb, BRO_K_LIT);
            return;
        }
        bro_parse_after_literals(d, ps, mb);
        return;
    }
    if (kind == BRO_K_CMD) {
        if (r != BRO_SYM_OK) { bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorInsertAndCopyLength : BRO_ST_UnexpectedEOF); return; }
        const uint32_t ie = bro_ic_insert[sym], ce = bro_ic_copy[sym];
        uint32_t insert_len = ie & 0xffffu, copy_len = ce & 0xffffu, extra;
        if (!bro_read_bits(d.in, ie >> 16, extra)) { bro_parse_finish(ps, BRO_ST_UnexpectedEOF); return; }   // insert extra bits first
        insert_len += extra;
        if (!bro_read_bits(d.in, ce >> 16, extra)) { bro_parse_finish(ps, BRO_ST_UnexpectedEOF); return; }
        copy_len += extra;
        if (ps.mlen < (d.pos - ps.mb_begin) + insert_len) { bro_parse_finish(ps, BRO_ST_ExceededExpectedBytes); return; }   // src/lib.rs:2036-2039
        ps.ins_rem = insert_len; ps.copy_len = copy_len; ps.need_dist = sym >= 128u;
        if (insert_len != 0u) bro_parse_enter(d, ps, mb, BRO_K_LIT);
        else bro_parse_after_literals(d, ps, mb);
        return;
    }
    if (kind == BRO_K_DIST || kind == BRO_K_COPY0) {
        if (r != BRO_SYM_OK) { bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_ParseErrorDistanceCode : BRO_ST_UnexpectedEOF); return; }
        uint32_t distance, max_allowed;
        int st = bro_resolve_distance(d, sym, mb.npostfix, mb.ndirect, distance, max_allowed);
        if (st) { bro_parse_finish(ps, st); return; }
        const uint32_t mb_out = d.pos - ps.mb_begin, copy_len = ps.copy_len;
        if (distance <= max_allowed) {
            // an LZ77 back-reference: phase two materialises it
            if (ps.mlen < mb_out + copy_len) { bro_parse_finish(ps, BRO_ST_ExceededExpectedBytes); return; }     // src/lib.rs:2105-2108
            if (copy_len > d.cap - d.pos) { bro_parse_finish(ps, BRO_ST_OutputTooSmall); return; }
            if (!bro_rec_push(d, d.pos, copy_len, BRO_REC_LZ, distance)) { bro_parse_finish(ps, BRO_ST_RecordsFull); return; }
            d.pos += copy_len;
        } else {
            // a static dictionary word (src/lib.rs:1506-1540): emitted here, it needs no earlier output
            if (copy_len < 4u || copy_len > 24u) { bro_parse_finish(ps, BRO_ST_InvalidLengthInStaticDictionary); return; }
            const uint32_t word_id = distance - max_allowed - 1u;
            const uint32_t bits = bro_dict_size_bits[copy_len];
            const uint32_t index = word_id & ((1u << bits) - 1u), tid = word_id >> bits;
            if (tid > 120u) { bro_parse_finish(ps, BRO_ST_InvalidTransformId); return; }
            const int n = bro_dict_word(*d.sc, d.dict, d.quirk_spec, copy_len, index, tid);
            if (n < 0) { bro_parse_finish(ps, BRO_ST_PanicUppercaseZero); return; }
            if (ps.mlen < mb_out + (uint32_t)n) { bro_parse_finish(ps, BRO_ST_ExceededExpectedBytes); return; }  // after the transform (Q10)
            if ((uint32_t)n > d.cap - d.pos) { bro_parse_finish(ps, BRO_ST_OutputTooSmall); return; }
            for (uint32_t i = 0; i < (uint32_t)n; i++) d.out[d.pos + i] = d.sc->word[i];
            d.pos += (uint32_t)n;
        }
        if (d.pos - ps.mb_begin == ps.mlen) { ps.kind = BRO_K_HEADER; return; }                                  // src/lib.rs:2128-2130
        bro_parse_enter(d, ps, mb, BRO_K_CMD);
        return;
    }
    const uint32_t c = ps.cat;
    if (kind == BRO_K_BTYPE) {
        if (r != BRO_SYM_OK) { bro_parse_finish(ps, r == BRO_SYM_HOLE ? BRO_ST_InvalidBlockSwitchCommandCode : BRO_ST_UnexpectedEOF); return; }
        BroBlockCat& bc = mb.cat[c];
        const uint32_t bt = sym == 0u ? bc.btype_prev : sym == 1u ? (bc.btype + 1u) % bc.nbl : sym - 2u;
        bc.btype_prev = bc.btype;       // committed now; a failing block count ends the stream anyway
        bc.btype = bt;
        ps.kind = BRO_K_BCOUNT; ps.toff = bc.t_count;
        return;
    }
    // BRO_K_BCOUNT
    if (r != BRO_SYM_OK) { bro_parse_finish(ps, BRO_ST_UnexpectedEOF); return; }
    if (sym > 25u) { bro_parse_finish(ps, BRO_ST_InvalidBlockCountCode); return; }
    const uint32_t be = bro_block_count[sym];
    uint32_t extra;
    if (!bro_read_bits(d.in, be >> 16, extra)) { bro_parse_finish(ps, BRO_ST_UnexpectedEOF); return; }
    const uint32_t bl = (be & 0xffffu) + extra - 1u;
    if (c == 0u) ps.blen0 = bl; else if (c == 1u) ps.blen1 = bl; else ps.blen2 = bl;
    ps.kind = ps.pending;
    ps.toff = bro_parse_table(d, ps, mb, ps.pending);
}
