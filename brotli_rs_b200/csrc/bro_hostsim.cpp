// Host simulation of the warp decoder with a 1-lane "warp" -- CPU TEST-SUITE ONLY.
// Built into tests/_build/libbro_hostsim.so by tests/hostsim.py; never part of libbrotli_b200.so.
// It lets the CPU test-suite fuzz the decoder logic of bro_decoder_core.h against the oracle without a GPU.
#define BRO_HOSTSIM 1
#include <stdlib.h>
#include <string.h>

#include "bro_decoder_core.h"

extern "C" const uint8_t bro_dictionary_blob[];

// the decoder's on-chip block (lane 3 of 32, with the device's interleave: BroTl) and the three root tables behind it
static void bro_hostsim_bind(BroDec& d, uint8_t* mem) {
    BroTl tl;
    tl.base = mem + 4u * 3u;
    bro_scratch_bind(d.scv, tl);
    d.in.ring = tl;
    uint16_t* roots = (uint16_t*)(mem + 32u * BRO_TL_BYTES);
    d.scv.root_lit = roots; d.scv.root_cmd = roots + 256; d.scv.root_dist = roots + 512;
}

// arena_u16 = 0 selects the worst-case arena of the warp kernel; BRO_THREAD_ARENA_U16 simulates the thread kernel
// (which may answer BRO_ST_ArenaTooSmall, upon which the product re-runs the stream with the warp kernel).
extern "C" int bro_hostsim_decode(const uint8_t* in, size_t in_len, uint8_t* out, size_t cap, size_t* out_len, int quirks,
                                  unsigned arena_u16) {
    if (arena_u16 == 0) arena_u16 = BRO_ARENA_U16_MAX;
    BroDec d;
    memset(&d, 0, sizeof(d));
    uint8_t* sc = (uint8_t*)calloc(32u * BRO_TL_BYTES + 3u * 256u * sizeof(uint16_t), 1);
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
    uint8_t* sc = (uint8_t*)calloc(32u * BRO_TL_BYTES + 3u * 256u * sizeof(uint16_t), 1);
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
