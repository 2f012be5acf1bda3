// Host simulation of PHASE TWO of the two-phase path -- CPU TEST-SUITE ONLY (tests/_build/libbro_hostsim.so, built by
// tests/hostsim.py; never part of libbrotli_b200.so).
//
// It executes a stream's copy records the way bro_copy_kernel (bro_kernels_copy.cu) does: 32 records at a time, split
// into groups of records none of which reads what a record of the group writes; a group of long records goes piece by
// piece through the VERY lane code of the kernel (bro_copy_piece.h: bro_piece_geo / bro_piece_load / bro_piece_store,
// run here lane by lane for G = 32, 16 or 8 lanes per piece, all loads of a step before its first store); a group of
// short records and a periodic fill are restated as byte loops (their kernel code is warp-wide).  What the CPU
// test-suite gets from this: the piece geometry and realignment of the kernel, and the claim that the records of a group
// may be executed in any order, checked against the oracle without a GPU.
#include <stdint.h>
#include <string.h>

#include "bro_copy_piece.h"
#include "bro_records.h"

#ifndef BRO_COPY_PIECES
#define BRO_COPY_PIECES 4      // as in bro_kernels_copy.cu
#endif

namespace {

struct Rec { uint32_t dst, len, kind, a; };

template <int G>
void run_pieces(uint8_t* out, const uint8_t* in, const Rec* r, const uint32_t* geo, const uint8_t* const* sp, uint32_t j, uint32_t e) {
    constexpr int PP = 32 / G, ROUNDS = (BRO_COPY_PIECES + PP - 1) / PP;
    (void)in;
    for (uint32_t k0 = j; k0 < e; k0 += (uint32_t)(PP * ROUNDS)) {
        static BroPieceData<G> D[ROUNDS][32];      // [round][lane]
        uint32_t m_geo[ROUNDS][32], m_dst[ROUNDS][32];
        for (int rd = 0; rd < ROUNDS; rd++)
            for (uint32_t lane = 0; lane < 32u; lane++) {
                m_geo[rd][lane] = 0;
                if (k0 + (uint32_t)(PP * rd) >= e) continue;
                const uint32_t bl = lane & (uint32_t)(G - 1), sub = lane / (uint32_t)G;
                const uint32_t k = k0 + (uint32_t)(PP * rd) + sub;
                const uint32_t ks = k & 31u;
                uint32_t g = geo[ks];
                if (PP > 1 && k >= e) g = 0;
                m_geo[rd][lane] = g;
                m_dst[rd][lane] = r[ks].dst;
                bro_piece_load<G>(D[rd][lane], sp[ks], g, bl);
            }
        for (int rd = 0; rd < ROUNDS; rd++)
            for (uint32_t lane = 0; lane < 32u; lane++) {
                if (k0 + (uint32_t)(PP * rd) >= e) continue;
                bro_piece_store<G>(D[rd][lane], out + m_dst[rd][lane], m_geo[rd][lane], lane & (uint32_t)(G - 1));
            }
    }
}

#ifndef BRO_COPY_QUADS
#define BRO_COPY_QUADS 2       // as in bro_kernels_copy.cu
#endif

// bro_run_pieces_staged of bro_kernels_copy.cu, trip by trip: issue step q (every lane), then consume step q - (NQ - 1)
template <int G>
void run_pieces_staged(uint8_t* out, const Rec* r, const uint32_t* geo, const uint8_t* const* sp, uint32_t j, uint32_t e) {
    constexpr int PP = 32 / G, NQ = BRO_COPY_QUADS;
    alignas(16) static uint8_t stage[NQ * PP * BRO_STAGE_SLOT_BYTES];
    const uint32_t nq = (e - j + (uint32_t)PP - 1u) / (uint32_t)PP;
    for (uint32_t q = 0; q < nq + (uint32_t)(NQ - 1); q++) {
        if (q < nq)
            for (uint32_t lane = 0; lane < 32u; lane++) {
                const uint32_t bl = lane & (uint32_t)(G - 1), sub = lane / (uint32_t)G;
                const uint32_t k = j + q * (uint32_t)PP + sub, ks = k & 31u;
                const uint32_t g = k >= e ? 0u : geo[ks];
                bro_piece_issue<G>(stage + ((q % (uint32_t)NQ) * (uint32_t)PP + sub) * BRO_STAGE_SLOT_BYTES, sp[ks], g, bl);
            }
        if (q + 1u < (uint32_t)NQ) continue;
        const uint32_t c = q - (uint32_t)(NQ - 1);
        for (uint32_t lane = 0; lane < 32u; lane++) {
            const uint32_t bl = lane & (uint32_t)(G - 1), sub = lane / (uint32_t)G;
            const uint32_t k = j + c * (uint32_t)PP + sub, ks = k & 31u;
            const uint32_t g = k >= e ? 0u : geo[ks];
            bro_piece_consume<G>(stage + ((c % (uint32_t)NQ) * (uint32_t)PP + sub) * BRO_STAGE_SLOT_BYTES, out + r[ks].dst, g, bl);
        }
    }
}

// bro_run_pieces_bulk of bro_kernels_copy.cu (the product): up to NS pieces of the group are fetched whole into their slots
// (one bulk copy each on the device: every granule that holds a source byte), then consumed one after the other by all 32 lanes
template <int NS>
void run_pieces_bulk(uint8_t* out, const Rec* r, const uint32_t* geo, const uint8_t* const* sp, uint32_t j, uint32_t e) {
    alignas(16) static uint8_t stage[NS * BRO_STAGE_SLOT_BYTES];
    for (uint32_t k0 = j; k0 < e; k0 += (uint32_t)NS) {
        const uint32_t cn = e - k0 < (uint32_t)NS ? e - k0 : (uint32_t)NS;
        for (uint32_t i = 0; i < cn; i++)
            for (uint32_t lane = 0; lane < 32u; lane++) bro_piece_issue<32>(stage + i * BRO_STAGE_SLOT_BYTES, sp[k0 + i], geo[k0 + i], lane);
        for (uint32_t i = 0; i < cn; i++)
            for (uint32_t lane = 0; lane < 32u; lane++) bro_piece_consume<32>(stage + i * BRO_STAGE_SLOT_BYTES, out + r[k0 + i].dst, geo[k0 + i], lane);
    }
}


// bro_run_pieces_win of bro_kernels_copy.cu (the product), step by step: the ring of the last BRO_WIN_BYTES of output, pieces
// whose source lies in it loaded from there (every lane of the step loads before any lane deposits), every piece deposited
// together with the few bytes phase one wrote in front of it.
#define BRO_WIN_NONE 0xffffffffu
struct WinState { uint32_t wlo, whi, seen, hit, tail_end; bool on; };
alignas(16) static uint8_t g_ring[BRO_WIN_BYTES];
static uint64_t g_win_hits, g_win_seen;     // pieces served by the ring / pieces moved by the window form (test-suite statistics)

template <int G>
void run_pieces_win(uint8_t* out, uint32_t out_mis, const Rec* r, const uint32_t* geo, const uint8_t* const* sp, const uint32_t* gapw,
                    uint32_t j, uint32_t e, bool dep, WinState& w) {
    constexpr int PP = 32 / G;
    uint32_t big = 0;
    for (uint32_t l = 0; l < 32u; l++) if (BRO_GEO_GAP(geo[l]) == BRO_GAP_BIG) big |= 1u << l;
    uint32_t k0 = j;
    while (k0 < e) {
        uint32_t m = e - k0 < (uint32_t)PP ? e - k0 : (uint32_t)PP;
        const uint32_t bm = (k0 < 31u ? big >> (k0 + 1u) : 0u) & ((1u << (m - 1u)) - 1u);
        if (bm) m = (uint32_t)__builtin_ffs((int)bm);
        const bool fresh = dep && (w.whi == BRO_WIN_NONE || ((big >> k0) & 1u));
        if (fresh) w.wlo = w.whi = r[k0].dst;
        static BroPieceData<G> D[32];
        uint32_t m_g[32], m_gn[32];
        bool m_in[32];
        for (uint32_t lane = 0; lane < 32u; lane++) {
            const uint32_t bl = lane & (uint32_t)(G - 1), sub = lane / (uint32_t)G;
            const uint32_t ks = (k0 + sub) & 31u;
            uint32_t g = geo[ks];
            if (sub >= m) g = 0;
            uint32_t gn = BRO_GEO_GAP(g);
            if (gn == BRO_GAP_BIG || (fresh && sub == 0u)) gn = 0;
            const uint32_t spos = (uint32_t)((uintptr_t)sp[ks] - (uintptr_t)out);
            const uint32_t len = BRO_GEO_HEAD(g) + 16u * BRO_GEO_NVEC(g) + BRO_GEO_TAIL(g);
            const bool inwin = dep && g != 0u && bro_win_holds(spos, len, w.wlo, w.whi);
            if (inwin) bro_piece_load_win<G>(D[lane], g_ring, spos + out_mis, g, bl);
            else bro_piece_load<G>(D[lane], sp[ks], g, bl);
            if (inwin && bl == 0u) { w.hit++; g_win_hits++; }
            m_g[lane] = g; m_gn[lane] = gn; m_in[lane] = inwin;
        }
        w.seen += m;
        g_win_seen += m;
        for (uint32_t lane = 0; lane < 32u; lane++) {
            const uint32_t bl = lane & (uint32_t)(G - 1), sub = lane / (uint32_t)G;
            const uint32_t ks = (k0 + sub) & 31u;
            const uint32_t g = m_g[lane], d = r[ks].dst;
            if (!dep) bro_piece_store<G>(D[lane], out + d, g, bl);
            else if (g != 0u) {
                bro_piece_store_win<G>(D[lane], out + d, g_ring, d + out_mis, g, bl);
                if (bl < m_gn[lane]) bro_stage_st8(g_ring + ((d + out_mis - m_gn[lane] + bl) & BRO_WIN_MASK), (gapw[ks] >> (8u * bl)) & 0xffu);
            }
        }
        (void)m_in;
        if (dep) w.whi = r[k0 + m - 1u].dst + r[k0 + m - 1u].len;
        k0 += m;
    }
}

}  // namespace

// -> pieces served by the ring / moved by the window form since the last call
extern "C" void bro_hostsim_copy_win_stats(uint64_t* hits_seen) { hits_seen[0] = g_win_hits; hits_seen[1] = g_win_seen; g_win_hits = g_win_seen = 0; }

// words: the records as phase one wrote them (4 x uint32 each).  group = lanes per piece (32, 16, 8, 4); + 100 = the staged
// form of the long-record path (BRO_COPY_STAGED) with that many lanes per piece; 200 + NS = the bulk form (BRO_COPY_BULK, the
// product of round 2's first half) with NS pieces in flight; 308 = the window form (BRO_COPY_WINDOW, the product): 8 lanes per
// piece, sources from the ring of the last BRO_WIN_BYTES of output where they lie in it.
// stats (optional, 3 words): groups executed by the piece path / as short records / periodic fills.
extern "C" void bro_hostsim_copy_exec(uint8_t* out, const uint8_t* in, const uint32_t* words, uint32_t nrec, int group, uint32_t* stats) {
    const uint32_t out_mis = (uint32_t)((uintptr_t)out & 15u);
    uint32_t st[3] = {0, 0, 0};
    WinState ws = {0, BRO_WIN_NONE, 0, 0, 0, true};
    for (uint32_t b = 0; b < nrec; b += 32u) {
        const uint32_t cnt = nrec - b < 32u ? nrec - b : 32u;
        Rec r[32];
        uint32_t geo[32];
        const uint8_t* sp[32];
        memset(r, 0, sizeof(r));
        for (uint32_t l = 0; l < cnt; l++) {
            const uint32_t* w = words + 4u * (b + l);
            r[l].dst = w[0]; r[l].len = w[1] & BRO_REC_LEN_MASK; r[l].kind = w[1] >> BRO_REC_KIND_SHIFT; r[l].a = w[2];
        }
        // (window form) the bytes phase one wrote between the record before and this one, as the kernel fetches them per batch
        uint32_t gapcode[32], gapw[32];
        memset(gapcode, 0, sizeof(gapcode)); memset(gapw, 0, sizeof(gapw));
        if (group >= 300 && ws.on) {
            for (uint32_t l = 0; l < cnt; l++) {
                const uint32_t prev = l ? r[l - 1u].dst + r[l - 1u].len : ws.tail_end;
                const uint32_t gap = r[l].dst - prev;
                gapcode[l] = gap <= 4u ? gap : BRO_GAP_BIG;
                if (gap - 1u < 4u) for (uint32_t i = 0; i < gap; i++) gapw[l] |= (uint32_t)out[prev + i] << (8u * i);
            }
            ws.tail_end = r[cnt - 1u].dst + r[cnt - 1u].len;
        }
        uint32_t j = 0;
        while (j < cnt) {
            uint32_t e = j;
            while (e < cnt && (r[e].kind == BRO_REC_STORED || (r[e].dst - r[e].a) + r[e].len <= r[j].dst)) e++;
            if (e == j) {      // a record that overlaps its own source: the kernel fills by doubling; byte order is the definition
                for (uint32_t i = 0; i < r[j].len; i++) out[r[j].dst + i] = out[r[j].dst + i - r[j].a];
                st[2]++;
                j++;
                ws.whi = BRO_WIN_NONE;
                continue;
            }
            uint32_t gsum = 0;
            bool too_long = false;
            for (uint32_t l = j; l < e; l++) {
                gsum += r[l].len;
                uint32_t head = (16u - ((r[l].dst + out_mis) & 15u)) & 15u;
                if (head > r[l].len) head = r[l].len;
                too_long |= ((r[l].len - head) >> 4) > BRO_REC_PIECE_VECS;
            }
            if (gsum >= 192u * (e - j) && !too_long) {
                for (uint32_t l = 0; l < 32u; l++) {
                    sp[l] = r[l].kind == BRO_REC_STORED ? in + r[l].a : (const uint8_t*)out + (r[l].dst - r[l].a);
                    geo[l] = bro_piece_geo(r[l].dst + out_mis, (uint32_t)(uintptr_t)sp[l], r[l].len);
                }
                if (group >= 300) {
                    bool stored = false;
                    for (uint32_t l = j; l < e; l++) stored |= r[l].kind == BRO_REC_STORED;
                    for (uint32_t l = 0; l < 32u; l++) geo[l] |= gapcode[l] << 28;
                    const bool dep = ws.on && !stored;
                    if (!dep) ws.whi = BRO_WIN_NONE;
                    run_pieces_win<8>(out, out_mis, r, geo, sp, gapw, j, e, dep, ws);
                    if (ws.seen >= 64u && 4u * ws.hit < ws.seen) ws.on = false;
                }
                else if (group == 208) run_pieces_bulk<8>(out, r, geo, sp, j, e);
                else if (group == 203) run_pieces_bulk<3>(out, r, geo, sp, j, e);
                else if (group == 108) run_pieces_staged<8>(out, r, geo, sp, j, e);
                else if (group == 116) run_pieces_staged<16>(out, r, geo, sp, j, e);
                else if (group == 132) run_pieces_staged<32>(out, r, geo, sp, j, e);
                else if (group == 4) run_pieces<4>(out, in, r, geo, sp, j, e);
                else if (group == 8) run_pieces<8>(out, in, r, geo, sp, j, e);
                else if (group == 16) run_pieces<16>(out, in, r, geo, sp, j, e);
                else run_pieces<32>(out, in, r, geo, sp, j, e);
                st[0]++;
            } else {
                // short records: the kernel's segmented copy moves the units of all records of the group in one sweep;
                // executing the records LAST TO FIRST checks that their order does not matter
                for (uint32_t l = e; l-- > j;) {
                    const uint8_t* s = r[l].kind == BRO_REC_STORED ? in + r[l].a : out + (r[l].dst - r[l].a);
                    memmove(out + r[l].dst, s, r[l].len);
                }
                st[1]++;
                ws.whi = BRO_WIN_NONE;
            }
            j = e;
        }
    }
    if (stats) { stats[0] = st[0]; stats[1] = st[1]; stats[2] = st[2]; }
}
