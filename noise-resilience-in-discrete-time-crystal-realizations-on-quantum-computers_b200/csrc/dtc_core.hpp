// dtc_core.hpp -- host-side program object: layer tables + pass schedule (no CUDA calls here).
// Shared by the CUDA library (dtcsim.cu) and the CPU emulation harness (tests/emul/emul.cpp).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "dtc_hd.cuh"
#include "dtc_stream.cuh"

struct DtcGenericStep {
    int kind;      // 0: rotation on qubit q of `layer`, 1: diagonal of `layer`
    int layer;
    int q;
    double t;      // tan(theta'/2)
};

struct DtcProgramHost {
    int n_qubits = 0, n_layers = 0, n_exec_layers = 0, n_local = 0, device = -1, engine = 0;
    int readout_bit = -1;          // hint (dtc_program_set_readout_hint): schedule the last pass on a tile holding this bit
    double global_phase = 0.0;
    bool finalized = false;
    int64_t n_sites = 0;
    std::vector<DtcEvent> events;
    std::vector<DtcLayer> layers;
    std::vector<DtcTilePass> passes;
    std::vector<DtcStreamPass> spasses;      // parallel to passes; mode == 0: not eligible for the streaming engine
    std::vector<DtcGenericStep> gsteps;
};

static inline int dtc_popc(u64 x) { return __builtin_popcountll(x); }

// A group of ten qubits starting at bit g is swept in tiles of 1024 runs of 64 B that span 2^(g-7) pages of 2 MiB; from
// g = 15 on that is more than the 128-entry TLB of an SM holds and the sweep becomes translation-bound (measured: 2.3 TB/s at
// g = 20 against 5.0 TB/s at g = 10).  Groups that start at or above this bit get five qubits and 2 KB runs instead.
#define DTC_HIGH_STRIDE_BIT (g_dtc_high_stride_bit)
static int g_dtc_high_stride_bit = 15;      // dtc_set_high_stride_bit(): tests lower it to reach this path at small n

// ---- raw event arrays -> staged DtcEvent list
static inline bool dtc_stage_events(DtcProgramHost& P, int64_t n, const int32_t* type, const int32_t* layer,
                                    const int32_t* q0, const int32_t* q1, const int32_t* slot, const double* val,
                                    const double* probs, double global_phase, std::string& err) {
    if (n < 0 || (n > 0 && (!type || !layer || !q0 || !q1 || !slot || !val || !probs))) {
        err = "NULL event array";
        return false;
    }
    P.events.resize((size_t)n);
    for (int64_t e = 0; e < n; ++e) {
        DtcEvent& E = P.events[(size_t)e];
        E.type = type[e]; E.layer = layer[e]; E.q0 = q0[e]; E.q1 = q1[e]; E.slot = slot[e]; E.k = 0;
        if (E.type == DTC_EVT_NOISE) {
            const double px = probs[3 * e], py = probs[3 * e + 1], pz = probs[3 * e + 2];
            if (px < 0 || py < 0 || pz < 0 || px + py + pz > 1.0 + 1e-12) {
                err = "invalid Pauli probabilities";
                return false;
            }
            E.c0 = px; E.c1 = px + py; E.c2 = px + py + pz;
        } else {
            E.c0 = val[e]; E.c1 = 0; E.c2 = 0;
        }
    }
    P.global_phase = global_phase;
    return true;
}

// ---- events -> layer tables ------------------------------------------------------------------
static inline bool dtc_build_layers(DtcProgramHost& P, std::string& err) {
    const int M = P.n_layers;
    P.layers.assign(M, DtcLayer());
    for (int j = 0; j < M; ++j) {
        DtcLayer& L = P.layers[j];
        memset(&L, 0, sizeof(L));
        L.cr = 1.0;
        for (int s = 0; s < 2; ++s)
            for (int q = 0; q < DTC_MAXQ; ++q) L.c1[s][q] = 1.0;
        for (int k = 0; k < DTC_MAXT; ++k) L.tc[k] = 1.0;
    }
    P.n_sites = 0;
    char buf[256];
    for (size_t e = 0; e < P.events.size(); ++e) {
        DtcEvent& E = P.events[e];
        if (E.q0 < 0 || E.q0 >= P.n_qubits || E.layer < 0 || E.layer >= M) {
            snprintf(buf, sizeof buf, "event %zu: qubit %d / layer %d out of range", e, E.q0, E.layer);
            err = buf;
            return false;
        }
        DtcLayer& L = P.layers[E.layer];
        if (E.type == DTC_EVT_ROT) {
            const double theta = E.c0;          // staged by set_events
            const double kk = nearbyint(theta / M_PI);
            const double thp = theta - kk * M_PI;
            long long k4 = ((long long)kk) % 4;
            if (k4 < 0) k4 += 4;
            E.k = (int)k4;
            const double t = tan(0.5 * thp);
            E.slot = (t == 0.0) ? 1 : 0;        // slot 1: pure Pauli RX(k pi), frame update only, no sign bit
            if (t != 0.0) {
                if (L.rot_any >> E.q0 & 1ull) {
                    snprintf(buf, sizeof buf, "event %zu: two rotations on qubit %d in layer %d", e, E.q0, E.layer);
                    err = buf;
                    return false;
                }
                L.rtan[E.q0] = t;
                L.rot_any |= 1ull << E.q0;
                L.cr *= cos(0.5 * thp);
            }
        } else if (E.type == DTC_EVT_D1) {
            if (E.slot < 0 || E.slot > 1 || (L.d1_any[E.slot] >> E.q0 & 1ull)) {
                snprintf(buf, sizeof buf, "event %zu: bad or duplicate D1 slot", e);
                err = buf;
                return false;
            }
            L.c1[E.slot][E.q0] = cos(0.5 * E.c0);
            L.s1[E.slot][E.q0] = sin(0.5 * E.c0);
            L.d1_any[E.slot] |= 1ull << E.q0;
        } else if (E.type == DTC_EVT_D2 || E.type == DTC_EVT_D2C) {
            if (E.slot < 0 || E.slot >= DTC_MAXT || E.q1 < 0 || E.q1 >= P.n_qubits || E.q1 == E.q0) {
                snprintf(buf, sizeof buf, "event %zu: bad D2 term", e);
                err = buf;
                return false;
            }
            L.tc[E.slot] = cos(0.5 * E.c0);
            L.ts[E.slot] = sin(0.5 * E.c0);
            L.ti[E.slot] = E.q0;
            L.tj[E.slot] = (E.type == DTC_EVT_D2C) ? DTC_VIRTUAL_QUBIT : E.q1;
            if (E.slot + 1 > L.n_terms) L.n_terms = E.slot + 1;
        } else if (E.type == DTC_EVT_NOISE) {
            P.n_sites++;
        } else {
            snprintf(buf, sizeof buf, "event %zu: unknown type %d", e, E.type);
            err = buf;
            return false;
        }
    }
    // global phase goes into the first diagonal layer's constant
    const double s = P.layers[0].cr;
    P.layers[0].cr = s * cos(P.global_phase);
    P.layers[0].ci = s * sin(P.global_phase);
    return true;
}

// ---- tile selection ----------------------------------------------------------------------------
// Choose 12 tile bits containing the active set S (<= 10 qubits) and global bits 0,1 such that S
// fits in a window of 10 consecutive tile-local positions starting at s2_lo in {0,1,2}.
static inline bool dtc_build_tile(u64 S, int n_local, int tb[DTC_TILE_BITS], int* s2_lo) {
    const u64 all = (n_local >= 64) ? ~0ull : ((1ull << n_local) - 1);
    if (dtc_popc(S) > 10 || (S & ~all)) return false;
    const u64 must = S | 3ull;
    const int need = DTC_TILE_BITS - dtc_popc(must);
    if (need < 0) return false;
    int top = -1;
    for (int b = 0; b < n_local; ++b)
        if ((S >> b) & 1ull) top = b;
    // preferred shapes (eligible for the TMA-fed streaming engine): [0,12) with the active set below bit 10,
    // or {0,1} + ten consecutive bits containing the active set
    // high-stride groups (lowest qubit >= DTC_HIGH_STRIDE_BIT, at most five qubits within five consecutive bits): seven
    // passive low bits + five consecutive bits -- 32 runs of 2 KB, so a tile touches at most 32 pages (mode C of the
    // streaming engine); ten active bits there would put each of 1024 runs of 64 B into a page of its own
    if (n_local >= DTC_TILE_BITS && top >= 0) {
        int low = 0;
        while (!((S >> low) & 1ull)) ++low;
        if (low >= DTC_HIGH_STRIDE_BIT && top - low < 5 && n_local >= 12) {
            int g = low;
            if (g + 5 > n_local) g = n_local - 5;
            for (int l = 0; l < 7; ++l) tb[l] = l;
            for (int l = 7; l < DTC_TILE_BITS; ++l) tb[l] = g + l - 7;
            *s2_lo = 2;
            return true;
        }
    }
    if (n_local >= DTC_TILE_BITS && top >= 0) {
        int g = -1;
        if (!(S >> 10)) g = 0;
        else if (!(S & 3ull)) {
            g = top + 1 - 10;
            if (g < 2) g = 2;
            if (g + 10 > n_local || (S & ((1ull << g) - 1))) g = -1;
        }
        if (g == 0 || g == 2) {
            for (int l = 0; l < DTC_TILE_BITS; ++l) tb[l] = l;
            *s2_lo = g;
            return true;
        }
        if (g > 2) {
            tb[0] = 0; tb[1] = 1;
            for (int l = 2; l < DTC_TILE_BITS; ++l) tb[l] = g + l - 2;
            *s2_lo = 2;
            return true;
        }
    }
    for (int klow = need; klow >= 0; --klow) {
        u64 tile = must;
        int added = 0;
        for (int b = 0; b < n_local && added < klow; ++b)          // lowest unused bits first
            if (!((tile >> b) & 1ull)) { tile |= 1ull << b; ++added; }
        for (int b = top + 1; b < n_local && added < need; ++b)     // then bits above the active set
            if (!((tile >> b) & 1ull)) { tile |= 1ull << b; ++added; }
        for (int b = top; b >= 0 && added < need; --b)              // then whatever is left below
            if (!((tile >> b) & 1ull)) { tile |= 1ull << b; ++added; }
        if (added < need) continue;
        int pos = 0, minpos = 99, maxpos = -1;
        for (int b = 0; b < n_local; ++b) {
            if (!((tile >> b) & 1ull)) continue;
            tb[pos] = b;
            if ((S >> b) & 1ull) {
                if (pos < minpos) minpos = pos;
                if (pos > maxpos) maxpos = pos;
            }
            ++pos;
        }
        if (maxpos < 0) { *s2_lo = 0; return true; }
        const int lo = (maxpos - 9 > 0) ? maxpos - 9 : 0;
        const int hi = (minpos < 2) ? minpos : 2;
        if (lo <= hi) { *s2_lo = lo; return true; }
    }
    return false;
}

static inline u64 dtc_lowest_bits(u64 m, int count) {
    u64 out = 0;
    while (m && count > 0) {
        const u64 b = m & (~m + 1);
        out |= b;
        m ^= b;
        --count;
    }
    return out;
}

static inline void dtc_classify_terms(DtcTilePass& T, const DtcLayer& L) {
    int loc[DTC_MAXQ];
    for (int q = 0; q < DTC_MAXQ; ++q) loc[q] = -1;
    for (int l = 0; l < DTC_TILE_BITS; ++l) loc[T.tb[l]] = l;
    const int S1 = T.s2_lo + 5;
    T.nT1 = T.nT2 = T.nX = T.nC = T.nO = 0;
    for (int k = 0; k < L.n_terms; ++k) {
        if (L.ts[k] == 0.0 && L.tc[k] == 1.0) continue;      // unused slot
        int a = loc[L.ti[k]], b = loc[L.tj[k]];
        if (a >= 0 && b >= 0) {
            if (a > b) { int t = a; a = b; b = t; }
            if (b <= S1) { T.T1k[T.nT1] = k; T.T1a[T.nT1] = a; T.T1b[T.nT1] = b; ++T.nT1; }
            else if (a >= S1) { T.T2k[T.nT2] = k; T.T2a[T.nT2] = a; T.T2b[T.nT2] = b; ++T.nT2; }
            else { T.Xk[T.nX] = k; T.Xa[T.nX] = a; T.Xb[T.nX] = b; ++T.nX; }
        } else if (a >= 0 || b >= 0) {
            T.Ck[T.nC] = k;
            T.Ca[T.nC] = (a >= 0) ? a : b;
            T.Cb[T.nC] = (a >= 0) ? L.tj[k] : L.ti[k];
            ++T.nC;
        } else {
            T.Ok[T.nO] = k; T.Oa[T.nO] = L.ti[k]; T.Ob[T.nO] = L.tj[k]; ++T.nO;
        }
    }
}

// ---- streaming engine eligibility ---------------------------------------------------------------
// A pass runs on k_tile_stream when its tile is [0,12) or {0,1} + [g,g+10), its active bits avoid the
// two passive positions of that layout, and every local two-body term fits one of the two tables.
static inline bool dtc_make_stream_pass(const DtcTilePass& T, const DtcLayer* LD, DtcStreamPass& S) {
    memset(&S, 0, sizeof(S));
    bool contig = true, modeB_shape = T.tb[0] == 0 && T.tb[1] == 1;
    for (int l = 0; l < DTC_TILE_BITS; ++l) contig = contig && T.tb[l] == l;
    for (int l = 3; l < DTC_TILE_BITS; ++l) modeB_shape = modeB_shape && T.tb[l] == T.tb[2] + (l - 2);
    unsigned active = 0;
    for (int l = 0; l < DTC_TILE_BITS; ++l)
        if (T.t1[l] != 0.0 || T.t2[l] != 0.0) active |= 1u << l;
    bool modeC_shape = true;
    for (int l = 0; l < 7; ++l) modeC_shape = modeC_shape && T.tb[l] == l;
    for (int l = 8; l < DTC_TILE_BITS; ++l) modeC_shape = modeC_shape && T.tb[l] == T.tb[7] + (l - 7);
    int mode = 0;
    if (modeC_shape && !contig && !(active & 0x7Fu)) mode = 3;
    else if (contig && !(active & 0xC00u)) mode = 1;
    else if ((contig || modeB_shape) && !(active & 0x3u)) mode = 2;
    if (!mode) return false;
    S.mode = mode;
    S.two = 2;
    S.contig = contig ? 1 : 0;
    S.n_local = T.n_local;
    S.g = mode == 3 ? T.tb[7] : T.tb[2];
    if (!contig && S.g + (mode == 3 ? 5 : 10) > T.n_local) return false;
    S.layerA = T.layerA; S.layerD = T.layerD; S.layerB = T.layerB;
    memcpy(S.tb, T.tb, sizeof(S.tb));
    memcpy(S.t1, T.t1, sizeof(S.t1));
    memcpy(S.t2, T.t2, sizeof(S.t2));
    S.tile_mask = T.tile_mask;
    if (T.layerD < 0 || !LD) return true;
    const DtcLayer& L = *LD;
    int loc[DTC_MAXQ];
    for (int q = 0; q < DTC_MAXQ; ++q) loc[q] = -1;
    for (int l = 0; l < DTC_TILE_BITS; ++l) loc[T.tb[l]] = l;
    struct Bond { int k, a, b, fam; };
    Bond bonds[DTC_MAXT];
    int nb = 0;
    for (int k = 0; k < L.n_terms; ++k) {
        if (L.ts[k] == 0.0 && L.tc[k] == 1.0) continue;      // unused slot
        int a = loc[L.ti[k]], b = loc[L.tj[k]];
        if (a >= 0 && b >= 0) {
            if (a > b) { int t = a; a = b; b = t; }
            int fam;
            if (mode == 3) {
                if (a >= 7) fam = 0;                         // core: both in [7,11]
                else if (b <= 2) fam = 3;
                else if (a >= 3 && b <= 6) fam = 4;
                else if (a <= 2 && b <= 6) fam = 5;
                else return false;                           // passive-active bond: register-fed kernel
            }
            else if (a >= 3 && b <= 7) fam = 0;
            else if (a == 2 && b == 3) fam = 1;
            else if (a == 7 && b == 8) fam = 2;
            else if (b <= 2) fam = 3;
            else if (a >= 8) fam = 4;
            else if (a <= 2 && b >= 8) fam = 5;
            else return false;                               // a bit of [3,7] bonded to a far bit: register-fed kernel
            bonds[nb++] = Bond{k, a, b, fam};
        } else if (a >= 0 || b >= 0) {
            S.Ck[S.nC] = k;
            S.Ca[S.nC] = (a >= 0) ? a : b;
            S.Cb[S.nC] = (a >= 0) ? L.tj[k] : L.ti[k];
            ++S.nC;
        } else {
            S.Ok[S.nO] = k; S.Oa[S.nO] = L.ti[k]; S.Ob[S.nO] = L.tj[k]; ++S.nO;
        }
    }
    int n = 0;
    for (int f = 0; f < 6; ++f) {
        S.fam_off[f] = n;
        for (int i = 0; i < nb; ++i)
            if (bonds[i].fam == f) { S.Fk[n] = bonds[i].k; S.Fa[n] = bonds[i].a; S.Fb[n] = bonds[i].b; ++n; }
    }
    S.fam_off[6] = n;
    return true;
}

// ---- pass schedule -----------------------------------------------------------------------------
// Qubits are partitioned once into groups of <= 10 (index order over the frequently rotated qubits;
// rarely rotated ones -- e.g. the Hadamard-test ancilla -- fill spare slots or get their own group).
// A pass works on one group G: it applies the remaining rotations of layer j on G; if that completes
// R_j it also applies D_j and, looking ahead, the rotations of R_{j+1} on G.  With two groups this is
// one state sweep per layer:  R_j|B D_j R_{j+1}|B  then  R_{j+1}|A D_{j+1} R_{j+2}|A  ...
struct DtcGroup {
    u64 members;
    int tb[DTC_TILE_BITS];
    int s2_lo;
};

static inline void dtc_add_group(std::vector<DtcGroup>& out, u64 members, int n_local) {
    if (!members) return;
    DtcGroup g;
    g.members = members;
    if (dtc_build_tile(members, n_local, g.tb, &g.s2_lo)) {
        out.push_back(g);
        return;
    }
    const int half = dtc_popc(members) / 2;           // cannot happen for a single qubit
    const u64 lo = dtc_lowest_bits(members, half);
    dtc_add_group(out, lo, n_local);
    dtc_add_group(out, members & ~lo, n_local);
}

static inline std::vector<DtcGroup> dtc_make_groups(const DtcProgramHost& P) {
    const int n = P.n_local;
    int count[DTC_MAXQ] = {0};
    int maxc = 0;
    for (int j = 0; j < P.n_exec_layers; ++j) {
        const DtcLayer& L = P.layers[j];
        for (int q = 0; q < n; ++q)
            if ((L.rot_any >> q) & 1ull) { ++count[q]; if (count[q] > maxc) maxc = count[q]; }
    }
    std::vector<u64> chunks;
    u64 cur = 0;
    int cap = 10, first = -1;
    for (int q = 0; q < n; ++q) {
        if (count[q] == 0 || count[q] * 2 <= maxc) continue;
        if (!cur) {
            first = q;
            cap = (q >= DTC_HIGH_STRIDE_BIT && n >= DTC_TILE_BITS) ? 5 : 10;
        } else if (cap == 5 && q - first >= 5) {            // five consecutive bits at most
            chunks.push_back(cur);
            cur = 0;
            first = q;
        }
        cur |= 1ull << q;
        if (dtc_popc(cur) == cap) { chunks.push_back(cur); cur = 0; }
    }
    if (cur) chunks.push_back(cur);
    u64 rare = 0;
    for (int q = 0; q < n; ++q) {
        if (count[q] == 0 || count[q] * 2 > maxc) continue;
        // nearest chunk with a free slot whose span would still fit a tile, else the "rare" group
        bool placed = false;
        for (size_t c = 0; c < chunks.size() && !placed; ++c) {
            if (dtc_popc(chunks[c]) >= 10) continue;
            int tb[DTC_TILE_BITS], s2;
            if (dtc_build_tile(chunks[c] | (1ull << q), n, tb, &s2)) { chunks[c] |= 1ull << q; placed = true; }
        }
        if (!placed) {
            rare |= 1ull << q;
            if (dtc_popc(rare) == 10) { chunks.push_back(rare); rare = 0; }
        }
    }
    if (rare) chunks.push_back(rare);
    std::vector<DtcGroup> groups;
    for (u64 c : chunks) dtc_add_group(groups, c, n);
    if (groups.empty()) dtc_add_group(groups, 1ull, n);      // circuit without rotations
    return groups;
}

// first_pick >= 0: group preferred for the first "diagonal only + look-ahead" pass (the passes then alternate between the
// groups, so this choice decides which group's tile the LAST pass works on -- see dtc_schedule_tile_for_readout).
static inline bool dtc_schedule_tile(DtcProgramHost& P, std::string& err, int first_pick = -1, int* n_groups = nullptr) {
    const int M = P.n_exec_layers, n = P.n_local;
    const u64 local_mask = (n >= 64) ? ~0ull : ((1ull << n) - 1);
    P.passes.clear();
    P.spasses.clear();
    for (int j = 0; j < M; ++j)
        if (P.layers[j].rot_any & ~local_mask) {
            err = "rotation on a non-local (global) qubit: exchange qubits before this layer";
            return false;
        }
    const std::vector<DtcGroup> groups = dtc_make_groups(P);
    if (n_groups) *n_groups = (int)groups.size();
    bool first_choice = true;
    int j = 0;
    u64 done = 0;
    while (j < M) {
        const DtcLayer& L = P.layers[j];
        const u64 rem = L.rot_any & ~done;
        const u64 next = (j + 1 < M) ? P.layers[j + 1].rot_any : 0ull;
        // a layer with nothing left to do (no rotation outstanding, identity diagonal) needs no sweep of its own:
        // its successor's rotations ride with that layer's diagonal instead (segments of a sharded run start so)
        // The last layer of a segment whose rotations all ran as the look-ahead of the previous pass and whose "diagonal" is
        // only the real normalisation constant of those tan-form rotations: the constant moves into the previous layer's
        // constant (everything is linear) instead of costing a sweep of its own.  Idempotent (the constant becomes 1).
        if (!rem && j == M - 1 && j >= 1 && L.n_terms == 0 && !L.d1_any[0] && !L.d1_any[1] && L.ci == 0.0 && L.cr != 1.0 &&
            !P.passes.empty() && P.passes.back().layerD == j - 1 && P.passes.back().layerB == j) {
            P.layers[j - 1].cr *= L.cr;
            P.layers[j - 1].ci *= L.cr;
            P.layers[j].cr = 1.0;
        }
        const bool trivial_d = L.n_terms == 0 && !L.d1_any[0] && !L.d1_any[1] && L.cr == 1.0 && L.ci == 0.0;
        if (!rem && trivial_d && (j + 1 < M || !P.passes.empty())) {
            // (the last layer too: its rotations were applied as the look-ahead of the previous pass and its diagonal is
            //  the identity -- segments of a sharded run end so)
            ++j;
            done = 0;
            continue;
        }
        // groups that still have work in this layer; the one kept for last gets D_j + look-ahead
        int pick = -1, n_cand = 0, best_keep = -1;
        for (size_t k = 0; k < groups.size(); ++k) {
            if (!(rem & groups[k].members)) continue;
            ++n_cand;
            const int ov = dtc_popc(next & groups[k].members);
            if (best_keep < 0 || ov > dtc_popc(next & groups[best_keep].members)) best_keep = (int)k;
        }
        bool complete;
        if (n_cand == 0) {                       // only D_j (+ look-ahead) left
            complete = true;
            for (size_t k = 0; k < groups.size(); ++k)
                if (pick < 0 || dtc_popc(next & groups[k].members) > dtc_popc(next & groups[pick].members)) pick = (int)k;
            if (first_choice && first_pick >= 0 && first_pick < (int)groups.size() &&
                dtc_popc(next & groups[first_pick].members) == dtc_popc(next & groups[pick].members))
                pick = first_pick;               // an equally good look-ahead group, chosen by the caller
            first_choice = false;
        } else if (n_cand == 1) {
            complete = true;
            pick = best_keep;
        } else {
            complete = false;
            for (size_t k = 0; k < groups.size(); ++k)
                if ((rem & groups[k].members) && (int)k != best_keep) { pick = (int)k; break; }
        }
        const DtcGroup& G = groups[pick];
        const u64 SA = rem & G.members;
        const u64 SB = complete ? (next & G.members) : 0ull;
        DtcTilePass T;
        memset(&T, 0, sizeof(T));
        memcpy(T.tb, G.tb, sizeof(T.tb));
        T.s2_lo = G.s2_lo;
        T.n_local = n;
        T.n_total = P.n_qubits;
        T.layerA = SA ? j : -1;
        T.layerD = (complete && !trivial_d) ? j : -1;      // an identity diagonal layer costs no tables and no multiplies
        T.layerB = (complete && SB) ? j + 1 : -1;
        for (int l = 0; l < DTC_TILE_BITS; ++l) {
            const int q = T.tb[l];
            T.t1[l] = ((SA >> q) & 1ull) ? L.rtan[q] : 0.0;
            T.t2[l] = ((SB >> q) & 1ull) ? P.layers[j + 1].rtan[q] : 0.0;
            if ((T.t1[l] != 0.0 || T.t2[l] != 0.0) && (l < T.s2_lo || l >= T.s2_lo + 10)) {
                err = "internal: active qubit outside the tile window";
                return false;
            }
        }
        if (T.layerD >= 0) dtc_classify_terms(T, L);
        for (int r = 0; r < DTC_NREG; ++r) {
            u64 o = 0;
            for (int k = 0; k < 5; ++k)
                if ((r >> k) & 1) o += 1ull << T.tb[T.s2_lo + 5 + k];
            T.roff[r] = o;
        }
        {
            const int S1 = T.s2_lo + 5;
            for (int b = 0; b < 7; ++b) T.tid_off[b] = 1ull << T.tb[b < S1 ? b : b + 5];
            u64 used = 0;
            for (int l = 0; l < DTC_TILE_BITS; ++l) used |= 1ull << T.tb[l];
            T.tile_mask = used;
            T.seg_n = 0;
            int src = 0, pos = 0;
            while (pos < n) {
                if ((used >> pos) & 1ull) { ++pos; continue; }
                int len = 0;
                while (pos + len < n && !((used >> (pos + len)) & 1ull)) ++len;
                if (T.seg_n >= 8) { err = "internal: too many tile index segments"; return false; }
                T.seg_src[T.seg_n] = src; T.seg_len[T.seg_n] = len; T.seg_dst[T.seg_n] = pos; ++T.seg_n;
                src += len; pos += len;
            }
        }
        P.passes.push_back(T);
        {
            DtcStreamPass S;
            if (!dtc_make_stream_pass(T, T.layerD >= 0 ? &L : nullptr, S)) S.mode = 0;
            P.spasses.push_back(S);
        }
        if (complete) {
            ++j;
            done = SB;
        } else {
            done |= SA;
        }
    }
    return true;
}

// Schedule such that the last pass works on a tile containing `bit` (the read-out qubit of a factorised circuit) when some
// choice of the first look-ahead group achieves that without adding passes: that pass can then reduce the qubit's
// density matrix itself instead of storing the state (k_tile_stream, fused read-out).
static inline bool dtc_schedule_tile_for_readout(DtcProgramHost& P, int bit, std::string& err) {
    int ng = 0;
    if (!dtc_schedule_tile(P, err, -1, &ng)) return false;
    auto last_has = [&]() {
        if (P.passes.empty()) return false;
        const DtcTilePass& T = P.passes.back();
        return ((T.tile_mask >> bit) & 1ull) && P.spasses.back().mode != 0;
    };
    if (bit < 0 || bit >= P.n_local || last_has()) return true;
    const size_t base_passes = P.passes.size();
    for (int g = 0; g < ng; ++g) {
        if (!dtc_schedule_tile(P, err, g)) return false;
        if (P.passes.size() <= base_passes && last_has()) return true;
    }
    return dtc_schedule_tile(P, err);             // no better choice: the default schedule
}

static inline void dtc_schedule_generic(DtcProgramHost& P) {
    P.gsteps.clear();
    for (int j = 0; j < P.n_exec_layers; ++j) {
        const DtcLayer& L = P.layers[j];
        for (int q = 0; q < P.n_qubits; ++q)
            if ((L.rot_any >> q) & 1ull) P.gsteps.push_back(DtcGenericStep{0, j, q, L.rtan[q]});
        const bool trivial = L.n_terms == 0 && !L.d1_any[0] && !L.d1_any[1] && L.cr == 1.0 && L.ci == 0.0;
        if (!trivial) P.gsteps.push_back(DtcGenericStep{1, j, -1, 0.0});
    }
}

static inline size_t dtc_workspace_bytes(const DtcProgramHost& P, int64_t n_traj) {
    size_t b = (size_t)(P.n_layers * 4 + 2) * (size_t)n_traj * sizeof(u64);
    b += (size_t)n_traj * sizeof(int);
    b = (b + 255) & ~(size_t)255;
    b += (size_t)n_traj * 4 * sizeof(double2);       // fused read-out: reduced density matrix of one qubit per trajectory
    b = (b + 255) & ~(size_t)255;
    // resident execution: one completion counter per (trajectory slot, pass), groups of at most 64 slots
    b += ((size_t)n_traj + 64) * (P.passes.empty() ? 1 : P.passes.size()) * sizeof(int);
    return (b + 255) & ~(size_t)255;
}

// offset of the completion counters of k_tile_resident inside the workspace
static inline size_t dtc_workspace_cnt_offset(const DtcProgramHost& P, int64_t n_traj) {
    size_t b = (size_t)(P.n_layers * 4 + 2) * (size_t)n_traj * sizeof(u64);
    b += (size_t)n_traj * sizeof(int);
    b = (b + 255) & ~(size_t)255;
    b += (size_t)n_traj * 4 * sizeof(double2);
    return (b + 255) & ~(size_t)255;
}

// offset of the fused-read-out density matrices inside the workspace
static inline size_t dtc_workspace_rdm_offset(const DtcProgramHost& P, int64_t n_traj) {
    size_t b = (size_t)(P.n_layers * 4 + 2) * (size_t)n_traj * sizeof(u64);
    b += (size_t)n_traj * sizeof(int);
    return (b + 255) & ~(size_t)255;
}
