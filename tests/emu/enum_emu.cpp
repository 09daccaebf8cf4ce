// enum_emu.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the configuration enumeration of k_expand
// (ambigram_b200/csrc/decode_common.cuh: lists, chunked cursor, column-mask restriction, light test) with g++ and checks
// it on random records against the flat decode of every configuration number (decode_light, the round-1/2 path that
// the warp-per-partition kernel still uses):
//   1. cursor_start at any number + cursor steps == decode_light of the same numbers (entry and pass bit);
//   2. the configurations that pass the light test under the restricted masks are exactly those that pass under the
//      full masks (restrict_masks removes only what the light test refuses one by one).
// The product never loads this (tests/test_enum_emu.py runs it).
#define AMBI_EMU 1
#include "simt_emu.h"
#include "../../ambigram_b200/csrc/decode_common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

using namespace ambi;

constexpr int NCOL = 32, GRP = 2;
struct SH {          // the fields the enumeration reads (GroupShared in stage_a.cu) + those of the flat decode
    uint8_t posP[GRP][NCOL], posE[GRP][NCOL];
    uint16_t L0[GRP][dec::L0CAP];
    uint16_t LE[GRP][dec::L0CAP + NCOL];
    uint32_t okP[GRP], parP[GRP], okL1[GRP], A2m[GRP], par1[GRP], A1m[GRP], tmask[GRP];
    uint8_t ovt[GRP][NCOL + 4];
    int32_t d0[GRP], f0[GRP], lslack[GRP];
    uint32_t nl0[GRP];
    uint8_t ptgs[GRP], ush[GRP];
};

struct Masks { uint32_t okP, parP, okL1, okL2, par1, par2; };

// both list forms of a record into slot `slot`; returns the number of configurations
static uint32_t stage(SH &S, int slot, const Masks &m)
{
    uint32_t nl0 = 0;
    const uint32_t total = dec::count_configs(m.okP, m.parP, m.okL1, m.okL2, m.par1, m.par2, &nl0);
    dec::write_lists(S, slot, m.okP, m.parP, m.okL1, m.okL2, m.par1, m.par2, nl0);     // L0, posE, par1 (flat decode)
    const uint32_t par1 = S.par1[slot];
    dec::write_lists2(S, slot, m.okP, m.parP, m.okL1, m.okL2, m.par1, m.par2, nl0);    // LE, A1m (cursor)
    S.par1[slot] = par1; S.nl0[slot] = nl0;
    return total;
}

extern "C" int emu_enum_check(uint64_t seed, int records, int scaled, long long *stats)
{
    std::mt19937_64 rng(seed);
    auto rnd = [&](uint32_t k) { return (uint32_t)(rng() % k); };
    static SH S;
    long long checked = 0, passed = 0, removed = 0, slow = 0, big = 0;
    for (int rec = 0; rec < records; ++rec) {
        memset(&S, 0, sizeof S);
        const int n = 2 + (int)rnd(NCOL - 1), a = (int)rnd(n);           // open columns are (a, n)
        const uint32_t open = (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) & ~((2u << a) - 1u);
        auto sub = [&](int dens) { uint32_t x = 0; for (int j = 0; j < 32; ++j) if ((int)rnd(100) < dens) x |= 1u << j; return x & open; };
        Masks m;
        // (every 16th record is dense: tens of thousands up to ~2.6e5 configurations, beyond 16 bits)
        const int dens = rec % 16 == 15 ? 97 + (int)rnd(4) : 20 + (int)rnd(70);
        m.okP = sub(dens); m.okL1 = sub(dens); m.okL2 = m.okL1 & sub(60);
        m.par2 = sub(rec % 16 == 15 ? 95 : 50); m.par1 = m.par2 | sub(50); m.parP = sub(rec % 16 == 15 ? 95 : 60);
        const uint32_t over = sub(40), fover = sub(40);
        const int ush = scaled ? (int)rnd(4) : 0;
        const int d0 = ((int)rnd(41) - 30) << ush, f0 = ((int)rnd(21) - 14) << ush, lslack = (int)rnd(3) ? (int)rnd(40 << ush) : (int)rnd(6 << ush);
        const uint32_t tm = rnd(4) == 0 ? sub(40) : 0u;
        const int ptgs = (int)rnd(6);
        for (int slot = 0; slot < GRP; ++slot) {
            dec::fill_ovt(S, slot, over, fover, n);
            S.d0[slot] = d0; S.f0[slot] = f0; S.lslack[slot] = lslack; S.tmask[slot] = tm; S.ptgs[slot] = (uint8_t)ptgs; S.ush[slot] = (uint8_t)ush;
        }
        // ---- 1. slot 0, full masks: flat decode of every number == cursor from every start
        const uint32_t total = stage(S, 0, m);
        const uint32_t stride = total > 60000 ? 1009u : 1u;        // the largest records are sampled
        if (total > 65535) ++big;
        if (S.nl0[0] > (uint32_t)dec::L0CAP) ++slow;
        std::vector<uint32_t> ent(total); std::vector<char> ok(total);
        std::vector<uint32_t> pass_full;
        for (uint32_t r = 0; r < total; ++r) {
            bool p = false;
            ent[r] = dec::decode_light(S, 0, r, tm, ush, &p);
            ok[r] = p && ent[r];
            if (ok[r]) pass_full.push_back(ent[r] & 0xffffffu);
        }
        for (uint32_t r0 = 0; r0 < total; r0 += stride) {
            dec::Cursor c;
            dec::cursor_start(S, 0, r0, c);
            const uint32_t steps = std::min<uint32_t>(5, total - r0);
            for (uint32_t k = 0; k < steps; ++k) {
                bool p = false;
                const uint32_t e = scaled ? dec::cursor_eval<true>(S, 0, c, S.nl0[0], &p) : dec::cursor_eval<false>(S, 0, c, S.nl0[0], &p);
                if (e != ent[r0 + k] || (bool)(p && e) != (bool)ok[r0 + k]) {
                    fprintf(stderr, "record %d: number %u from start %u: cursor %08x/%d, flat decode %08x/%d\n", rec, r0 + k, r0, e, (int)p, ent[r0 + k], (int)ok[r0 + k]);
                    return 1;
                }
                dec::cursor_advance(S, 0, c, S.nl0[0]);
                ++checked;
            }
        }
        // ---- 2. slot 1, restricted masks: the same set of passing configurations
        Masks r = m;
        std::vector<uint32_t> pass_res;
        if (lslack >= 0) {
            if (dec::restrict_masks(over, fover, lslack, d0, f0, ush, &r.okP, &r.okL1, &r.okL2)) ++removed;
            const uint32_t t2 = stage(S, 1, r);
            for (uint32_t q = 0; q < t2; ++q) {
                bool p = false;
                const uint32_t e = dec::decode_light(S, 1, q, tm, ush, &p);
                if (p && e) pass_res.push_back(e & 0xffffffu);
            }
            if (t2 > total) { fprintf(stderr, "record %d: restriction grew the space\n", rec); return 2; }
        }
        std::sort(pass_full.begin(), pass_full.end()); std::sort(pass_res.begin(), pass_res.end());
        if (pass_full != pass_res) {
            fprintf(stderr, "record %d: %zu configurations pass under the full masks, %zu under the restricted ones\n", rec, pass_full.size(), pass_res.size());
            return 3;
        }
        passed += (long long)pass_full.size();
    }
    stats[0] = checked; stats[1] = passed; stats[2] = removed; stats[3] = slow; stats[4] = big;
    return 0;
}
