// Host simulation of the TWO-PHASE path -- CPU TEST-SUITE ONLY (tests/_build/libbro_hostsim_parse.so, built by
// tests/hostsim.py; never part of libbrotli_b200.so).  Phase one is the very code the parse kernel runs per lane
// (bro_parse.h); phase two is replaced by the obvious byte loop over the copy records, so
// that the CPU test-suite can check the records phase one writes against the oracle without a GPU.
#define BRO_HOSTSIM 1
#define BRO_PARSE 1
#include <stdlib.h>
#include <string.h>

#include "bro_parse.h"

extern "C" const uint8_t bro_dictionary_blob[];

// Returns the status phase one leaves (BRO_ST_NeedFused etc. included).  *n_rec = records written, *n_steps = rounds of
// the machine.  rec_cap = 0 selects the product's share: one record per 2 compressed bytes + 32.
static uint32_t g_sizing = 0;         // next decodes only measure (bro_batch_sizes' mode): no output, unbounded slot
extern "C" void bro_hostsim_parse_set_sizing(unsigned on) { g_sizing = on; }

// phase two of the next decodes: 0 = the obvious byte loop over the records; 32, 16, 8 = the copy kernel's grouping and
// piece code with that many lanes per piece (bro_hostsim_copy.cpp); -1 = none (phase one only)
static int g_copy_group = 0;
static uint32_t g_copy_stats[3];
extern "C" void bro_hostsim_parse_set_copy_group(int group) { g_copy_group = group; }
extern "C" void bro_hostsim_parse_copy_stats(uint32_t* stats3) { memcpy(stats3, g_copy_stats, sizeof(g_copy_stats)); }
extern "C" void bro_hostsim_copy_exec(uint8_t* out, const uint8_t* in, const uint32_t* words, uint32_t nrec, int group, uint32_t* stats);

// look-ups through the narrow insert&copy / distance roots since the last call: [cmd all, 8-bit root, search, dist all, ...]
extern "C" void bro_hostsim_parse_root_stats(uint64_t* six) {
    memcpy(six, bro_hostsim_root_stats, sizeof(bro_hostsim_root_stats));
    memset(bro_hostsim_root_stats, 0, sizeof(bro_hostsim_root_stats));
}

static uint32_t* g_rec_out = 0;       // when set: the records of the next decode are copied here (4 words each)
static unsigned g_rec_out_cap = 0;

// Ask the next bro_hostsim_parse_decode to export its records (tests of the piece geometry).
extern "C" void bro_hostsim_parse_export_records(uint32_t* words, unsigned max_records) { g_rec_out = words; g_rec_out_cap = max_records; }

extern "C" int bro_hostsim_parse_decode(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap, size_t* out_len,
                                        int quirks, unsigned arena_u16, unsigned rec_cap, unsigned* n_rec, unsigned* n_steps) {
    if (arena_u16 == 0) arena_u16 = BRO_THREAD_ARENA_U16;
    if (rec_cap == 0) rec_cap = (unsigned)(in_len >> 1) + 32u;
    BroDec d;
    memset(&d, 0, sizeof(d));
    uint16_t* arena = (uint16_t*)malloc(2u * (size_t)arena_u16);
    BroRec* rec = (BroRec*)malloc(sizeof(BroRec) * (size_t)rec_cap);
    // the thread's on-chip block, with the device's interleave: lane 7 of 32 (bro_decoder_core.h, BroTl)
    uint8_t* blocks = (uint8_t*)calloc(32u * BRO_TL_BYTES, 1);
    {
        BroTl tl;
        tl.base = blocks + 4u * 7u;
        bro_scratch_bind(d.scv, tl);
        d.in.ring = tl;
    }
    d.arena = arena;
    d.arena_cap = arena_u16;
    d.arena_base = 0;
    d.dict = bro_dictionary_blob;
    d.out = out;
    d.out_mis = (uint32_t)((uintptr_t)out & 15u);
    d.sizing = g_sizing;
    d.cap = (cap > BRO_MAX_SLOT || g_sizing) ? (uint32_t)BRO_MAX_SLOT : (uint32_t)cap;
    d.pos = 0;
    d.p1 = d.p2 = 0;
    d.d0 = 4; d.d1 = 11; d.d2 = 15; d.d3 = 16;
    d.quirk_spec = quirks;
    static uint32_t ic[48];
    for (unsigned k = 0; k < 48; k++) bro_ic_compact_entry(k, ic[k]);
    for (unsigned i = 0; i < 704; i++) {       // the compact form answers exactly what the reference's table holds
        uint32_t ie, ce;
        bro_ic_lookup(ic, i, ie, ce);
        if (ie != bro_ic_insert[i] || ce != bro_ic_copy[i]) abort();
    }
    d.rec = rec; d.nrec = 0; d.rec_cap = rec_cap; d.in_base = in; d.ic = ic;
    bro_bits_init(d.in, in, in + in_len);
    BroParse ps;
    BroMbInfo mb;
    memset(&mb, 0, sizeof(mb));
    bro_parse_begin(ps);
    unsigned steps = 0;
    while (ps.kind != BRO_K_DONE) {
        if (ps.kind == BRO_K_HEADER) bro_parse_header(d, ps, mb);
        else { if (d.imm) bro_parse_round<true>(d, ps, mb); else bro_parse_round<false>(d, ps, mb); steps++; }
    }
    // phase two, the obvious way
    if (g_copy_group < 0) {
        // phase one only: the caller executes the exported records itself (bro_warpsim_copy.cpp: the copy kernel's own code)
    } else if (ps.st == BRO_ST_OK && !g_sizing && g_copy_group) {
        static_assert(sizeof(BroRec) == 16, "a record is four words");
        bro_hostsim_copy_exec(out, in, (const uint32_t*)rec, d.nrec, g_copy_group, g_copy_stats);
    } else if (ps.st == BRO_ST_OK && !g_sizing) {
        for (uint32_t k = 0; k < d.nrec; k++) {
            const BroRec r = rec[k];
            const uint32_t len = r.len_kind & BRO_REC_LEN_MASK, kind = r.len_kind >> BRO_REC_KIND_SHIFT;
            if (kind == BRO_REC_STORED) memcpy(out + r.dst, in + r.a, len);
            else for (uint32_t i = 0; i < len; i++) out[r.dst + i] = out[r.dst + i - r.a];
        }
    }
    if (g_rec_out) {
        for (uint32_t k = 0; k < d.nrec && k < g_rec_out_cap; k++) {
            g_rec_out[4 * k] = rec[k].dst; g_rec_out[4 * k + 1] = rec[k].len_kind; g_rec_out[4 * k + 2] = rec[k].a; g_rec_out[4 * k + 3] = d.out_mis;
        }
        g_rec_out = 0;
    }
    *out_len = d.pos;
    if (n_rec) *n_rec = d.nrec;
    if (n_steps) *n_steps = steps;
    free(rec);
    free(arena);
    free(blocks);
    return ps.st;
}
