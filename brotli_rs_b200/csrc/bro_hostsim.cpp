// Host simulation of the warp decoder with a 1-lane "warp" -- CPU TEST-SUITE ONLY.
// Built into tests/_build/libbro_hostsim.so by tests/hostsim.py; never part of libbrotli_b200.so.
// It lets the CPU test-suite fuzz the decoder logic of bro_decoder_core.h against the oracle without a GPU.
#define BRO_HOSTSIM 1
#include <stdlib.h>
#include <string.h>

// test-suite instrumentation: what the meta-block headers of the last decodes announced (histograms, saturating at 63)
static unsigned g_mb_hist[5][64];
#define BRO_MB_STAT(ntl, ntd, n0, n1, n2, st) do { if (!(st)) { unsigned v_[5] = {(ntl), (ntd), (n0), (n1), (n2)}; \
    for (int q_ = 0; q_ < 5; q_++) g_mb_hist[q_][v_[q_] < 63u ? v_[q_] : 63u]++; } } while (0)
#include "bro_decoder_core.h"
extern "C" void bro_hostsim_mb_hist(unsigned* out, int reset) { memcpy(out, g_mb_hist, sizeof(g_mb_hist)); if (reset) memset(g_mb_hist, 0, sizeof(g_mb_hist)); }

extern "C" const uint8_t bro_dictionary_blob[];

// the decoder's on-chip block (lane 3 of 32, with the device's interleave: BroTl) and the three root tables behind it
static void bro_hostsim_bind(BroDec& d, uint8_t* mem) {
    BroTl tl;
    tl.base = mem + 4u * 3u;
    bro_scratch_bind(d.scv, tl);
    d.in.ring = tl;
    uint16_t* roots = (uint16_t*)(mem + 32u * BRO_TL_BYTES);
    d.scv.root_lit = roots; d.scv.root_cmd = roots + 256; d.scv.root_dist = roots + 512;
    d.scv.hot_lit = roots + 768;
    d.scv.hot_cmap_l = (uint8_t*)(roots + 768 + BRO_HOT_LIT_U16);
    d.scv.hot_cmap_d = d.scv.hot_cmap_l + BRO_HOT_CMAP_L;
    d.scv.hot_modes = d.scv.hot_cmap_d + BRO_HOT_CMAP_D;
}
#define BRO_HOSTSIM_SC_BYTES (32u * BRO_TL_BYTES + (768u + BRO_HOT_LIT_U16) * sizeof(uint16_t) + BRO_HOT_CMAP_L + BRO_HOT_CMAP_D + BRO_HOT_MODES)

// arena_u16 = 0 selects the worst-case arena of the warp kernel; BRO_THREAD_ARENA_U16 simulates the thread kernel
// (which may answer BRO_ST_ArenaTooSmall, upon which the product re-runs the stream with the warp kernel).
extern "C" int bro_hostsim_decode(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap, size_t* out_len, int quirks,
                                  unsigned arena_u16) {
    if (arena_u16 == 0) arena_u16 = BRO_ARENA_U16_MAX;
    BroDec d;
    memset(&d, 0, sizeof(d));
    uint8_t* sc = (uint8_t*)calloc(BRO_HOSTSIM_SC_BYTES, 1);
    uint16_t* arena = (uint16_t*)malloc(2u * (size_t)arena_u16);
    bro_hostsim_bind(d, sc);
    d.arena = arena;
    d.arena_cap = arena_u16;
    d.arena_base = 0;
    d.dict = bro_dictionary_blob;
    d.out = out;
    d.cap = cap > BRO_MAX_SLOT ? (uint32_t)BRO_MAX_SLOT : (uint32_t)cap;
    d.pos = 0;
    d.p1 = d.p2 = 0;
    d.d0 = 4; d.d1 = 11; d.d2 = 15; d.d3 = 16;
    d.quirk_spec = quirks;
    bro_bits_init(d.in, in, in + in_len);
    int st = bro_decode_stream(d);
    *out_len = d.pos;
    free(sc);
    free(arena);
    return st;
}

extern "C" unsigned bro_hostsim_thread_arena_u16() { return BRO_THREAD_ARENA_U16; }

// The resumable decode (bro_decode_stream_resume, the device side of the streaming reader): one call over the input and
// output the caller has at hand, from and to the resume point *ck.  out[0 .. ck->pos) must hold the history (the last
// min(window, bytes so far) bytes of output).  *out_len = bytes in the slot when the call stopped (only those in front
// of the new ck->pos are final).
extern "C" int bro_hostsim_decode_resume(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap, size_t* out_len, int quirks,
                                         BroResume* ck) {
    BroDec d;
    memset(&d, 0, sizeof(d));
    uint8_t* sc = (uint8_t*)calloc(BRO_HOSTSIM_SC_BYTES, 1);
    uint16_t* arena = (uint16_t*)malloc(2u * (size_t)BRO_ARENA_U16_MAX);
    bro_hostsim_bind(d, sc);
    d.arena = arena;
    d.arena_cap = BRO_ARENA_U16_MAX;
    d.arena_base = 0;
    d.dict = bro_dictionary_blob;
    d.out = out;
    d.cap = cap > BRO_MAX_SLOT ? (uint32_t)BRO_MAX_SLOT : (uint32_t)cap;
    d.quirk_spec = quirks;
    int st = bro_decode_stream_resume(d, ck, in, in + in_len);
    *out_len = d.pos;
    free(sc);
    free(arena);
    return st;
}
