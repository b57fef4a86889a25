// dtc_stream.cuh -- per-thread code of the TMA-fed streaming tile engine (k_tile_stream in dtcsim.cu).
//
// Same fused pass as the register-fed k_tile_pass ( R_A|G -> D -> R_B|G  over tiles of 2^12 amplitudes;
// reference hot path: the per-shot statevector evolution inside AerSimulator.run(), fast.py:211), but
// the tile lives in shared memory from the moment the TMA engine lands it until the TMA engine stores
// it, so loads and stores occupy neither registers nor issue slots of the compute warps and several
// tiles per SM are in flight.  Compiles under nvcc and under g++ (tests/emul/emul.cpp runs these
// functions thread by thread on the CPU).
//
// Tile-local index l (12 bits) = 16 B chunk index in the stage buffer (dense, no swizzle).  Two layouts:
//   mode A  tile = local qubits 0..11 (64 KB contiguous);  rotatable bits 0..9,  passive 10,11
//   mode B  tile = qubits {0,1} + [g, g+10) (1024 runs of 64 B);  rotatable bits 2..11, passive 0,1
//   mode C  tile = qubits 0..6 + [g, g+5) (32 runs of 2 KB);  rotatable bits 7..11, passive 0..6.  For groups whose
//           64 B runs would each lie in a different 2 MiB page (g >= 15: translation-bound); one register set, one phase.
// Register sets:  S2 = bits 3..7 in modes A, B;  S1 = {0,1,2,8,9} (A) / {2,8,9,10,11} (B);  mode C: bits 7..11 only.
// Bank conflicts: a 16 B access is conflict free when the 8 lanes of a quarter warp hit 8 different
// values of l[0:2].  Phase 2 maps those lanes to l[0:2] directly.  Phases 1/3 hold low bits in
// registers, so lane c reads its registers in the order l_low = k ^ c: the tan-form RX butterfly is
// symmetric under exchanging the two members of a pair, so the permuted order needs no fix-up.
#pragma once
#include "dtc_hd.cuh"

#if defined(__CUDACC__)
#define DTC_NOUNROLL _Pragma("unroll 1")
#else
#define DTC_NOUNROLL
#endif

#define DTC_STREAM_STAGES 3
#define DTC_STREAM_WG 2                       // compute warpgroups (128 threads each) per CTA
#define DTC_STREAM_THREADS (128 * DTC_STREAM_WG + 32 + 32 * DTC_STREAM_STAGES)   // + TMA driver warp + one table-builder warp per stage

struct DtcStreamPass {
    int mode;                      // 1: A, 2: B, 3: C
    int contig;                    // tile is 64 KB contiguous in global memory (bulk copy, no tensor map)
    int n_local, g;                // B: tile bits 2..11 = global bits g..g+9
    int layerA, layerD, layerB;
    int two;                       // = 2: opaque trip count that keeps shared code blocks rolled (instruction footprint)
    int tmap_slot;                 // k_tile_resident: which of its two tensor maps this pass uses (non-contiguous tiles)
    int tb[DTC_TILE_BITS];
    double t1[DTC_TILE_BITS], t2[DTC_TILE_BITS];
    u64 tile_mask;
    // two-body terms of D_layerD, local-local ones by table family:
    //   0: core  both ends in local [3,7]      1: (2,3)                    2: (7,8)
    //   3: T2lo  both in {0,1,2}               4: T2hi  both in [8,11]     5: T2x  one in {0,1,2}, one in [8,11]
    // (mode C: core = [7,11], T2hi = [3,6], families 1 and 2 empty)
    int fam_off[7];                // family f = entries [fam_off[f], fam_off[f+1]) of Fk/Fa/Fb
    int nC, nO;
    unsigned char Fk[DTC_MAXT], Fa[DTC_MAXT], Fb[DTC_MAXT];
    unsigned char Ck[DTC_MAXT], Ca[DTC_MAXT], Cb[DTC_MAXT];      // local a, outer qubit b
    unsigned char Ok[DTC_MAXT], Oa[DTC_MAXT], Ob[DTC_MAXT];      // outer, outer
} __attribute__((aligned(16)));

// What the TMA driver and the compute warps of k_tile_resident need of a pass: small enough for an array of these to travel
// in the kernel's parameter space (read through the constant cache, like the single grid-constant pass of k_tile_stream).
#define DTC_RESIDENT_MAX_PASSES 128
struct DtcResidentPass {
    unsigned char mode, contig, g, tmap_slot;
    short layerA, layerD, layerB, n_local;
    unsigned char tb[DTC_TILE_BITS];
    double t1[DTC_TILE_BITS], t2[DTC_TILE_BITS];
    u64 tile_mask;
};

// per-stage phase tables of the tile in that stage, written by the stage's table-builder warp.
// Phase of tile-local index l = T1c[l[3:7]] * F[l2][l3] * G[l8][l7] * T2[l[0:2], l[8:11]]
struct StreamSlot {
    double2 T1c[32];               // one-body of local 3..7 and bonds inside [3,7]; index = l[3:7] (the phase-2 register)
    double2 F[2][2], G[2][2];      // bonds (2,3): F[l2][l3];  bonds (7,8): G[l8][l7]
    double2 T2[128];               // index = local bits 0,1,2,8,9,10,11 (= thread id of phase 2), times the tile constant
    u64 rmA, rmB;                  // rotation sign masks of the tile's trajectory
};

// scratch of one table-builder warp
struct StreamBuild {
    double2 E[DTC_TILE_BITS][2];
    double2 B[DTC_MAXT][2];
    double2 T1c[32], FG[8], T2lo[8], T2hi[16];
    double2 scratch[32];
    double2 C;
};

// ---- tile geometry
template <class PassT>
DTC_HD u64 stream_tile_base(u64 tile_in_traj, const PassT& P) {
    if (P.contig) return tile_in_traj << DTC_TILE_BITS;
    if (P.mode == 3) {
        const int lb = P.g - 7;
        return ((tile_in_traj & ((1ull << lb) - 1)) << 7) | ((tile_in_traj >> lb) << (P.g + 5));
    }
    const int lb = P.g - 2;
    return ((tile_in_traj & ((1ull << lb) - 1)) << 2) | ((tile_in_traj >> lb) << (P.g + 10));
}

// first local bit of the core table (register set of the diagonal phase) and of the high half of T2
DTC_HD int stream_core_base(int mode) { return mode == 3 ? 7 : 3; }
DTC_HD int stream_hi_base(int mode) { return mode == 3 ? 3 : 8; }

// chunk index of register k in phases 1/3
template <int MODE>
DTC_HD int stream_chunk13(int t, int k) {
    const int lane = t & 31, w = t >> 5, c = lane & 7, h = lane >> 3;
    if (MODE == 1) return ((k & 7) ^ c) | (c << 3) | (h << 6) | ((k >> 3) << 8) | (w << 10);
    const int c2 = c >> 2;
    return (c & 3) | (((k & 1) ^ c2) << 2) | (c2 << 3) | (h << 4) | (w << 6) | ((k >> 1) << 8);
}
// ... in phase 2 (register r <-> local bits 3..7)
DTC_HD int stream_chunk2(int t, int r) {
    const int lane = t & 31, w = t >> 5;
    return (lane & 7) | (r << 3) | ((lane >> 3) << 8) | (w << 10);
}

// local bit rotated by register bit j in phases 1/3
template <int MODE>
DTC_HD int stream_s1_bit(int j) {
    if (MODE == 1) return j < 3 ? j : j + 5;
    return j == 0 ? 2 : j + 7;
}

template <int MODE, class TB>
DTC_HD void stream_signed_s1(const double* tbase, const TB* tb, u64 rmask, double out[5]) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int l = stream_s1_bit<MODE>(j);
        const double t = tbase[l];
        out[j] = ((rmask >> tb[l]) & 1ull) ? -t : t;
    }
}

// ---- phases (tile: the stage buffer, 4096 chunks of 16 B)
template <int MODE>
DTC_HD void stream_phase13_signed(int t, double2* tile, const double tt[5]) {
    double2 a[DTC_NREG];
    // MODE A: 8 address registers (k & 7) + immediates; MODE B: 2 + immediates
    double2* p[8];
    constexpr int NP = (MODE == 1) ? 8 : 2;
#pragma unroll
    for (int j = 0; j < NP; ++j) p[j] = tile + stream_chunk13<MODE>(t, j);
#pragma unroll
    for (int k = 0; k < DTC_NREG; ++k) a[k] = p[k & (NP - 1)][(MODE == 1) ? ((k >> 3) << 8) : ((k >> 1) << 8)];
    tile_rot5(a, tt);
#pragma unroll
    for (int k = 0; k < DTC_NREG; ++k) p[k & (NP - 1)][(MODE == 1) ? ((k >> 3) << 8) : ((k >> 1) << 8)] = a[k];
}

template <int MODE>
DTC_HD void stream_phase13(int t, double2* tile, const double* tbase, const int* tb, u64 rmask) {
    double tt[5];
    stream_signed_s1<MODE>(tbase, tb, rmask, tt);
    stream_phase13_signed<MODE>(t, tile, tt);
}

template <class PassT>
DTC_HD void stream_phase2(int t, double2* tile, const StreamSlot& tab, const PassT& P, u64 rmA, u64 rmB) {
    double tA[5], tB[5];
    tile_signed_t(P.t1, P.tb, 3, rmA, tA);
    tile_signed_t(P.t2, P.tb, 3, rmB, tB);
    double2 a[DTC_NREG];
    double2* p = tile + stream_chunk2(t, 0);
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) a[r] = p[r << 3];
    // thread constants c[l7][l3] = T2[t] * F[l2][l3] * G[l8][l7]   (l2 = lane bit 2, l8 = lane bit 3)
    const int lane = t & 31, l2 = (lane >> 2) & 1, l8 = (lane >> 3) & 1;
    const double2 cf0 = cmul(tab.T2[t], tab.F[l2][0]), cf1 = cmul(tab.T2[t], tab.F[l2][1]);
    const double2 c00 = cmul(cf0, tab.G[l8][0]), c01 = cmul(cf1, tab.G[l8][0]);
    const double2 c10 = cmul(cf0, tab.G[l8][1]), c11 = cmul(cf1, tab.G[l8][1]);
    // (a rolled two-iteration loop sharing the two four-level blocks was measured slower: selects + spills)
    tile_rot_bits(a, tA, 0, 4);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        rot_pair(a[i], a[i | 16], tA[4]);
        // T1c is read at warp-uniform addresses (one shared-memory wavefront per load)
        const double2 p0 = cmul(tab.T1c[i], (i & 1) ? c01 : c00);
        const double2 p1 = cmul(tab.T1c[i | 16], (i & 1) ? c11 : c10);
        a[i] = cmul(a[i], p0);
        a[i | 16] = cmul(a[i | 16], p1);
        rot_pair(a[i], a[i | 16], tB[4]);
        DTC_SCHED_FENCE();
    }
    tile_rot_bits(a, tB, 0, 4);
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) p[r << 3] = a[r];
}

// Phase 2 of a pass that lacks one of its three parts (first / last pass of a circuit, rotation-only sweeps of large
// registers): the absent parts are skipped instead of being executed with zero angles / identity tables.  Kept out of
// stream_phase2 so that the register allocation of the hot path (all three parts present) is not disturbed.
template <class PassT>
DTC_HD void stream_phase2_partial(int t, double2* tile, const StreamSlot& tab, const PassT& P, u64 rmA, u64 rmB) {
    double2 a[DTC_NREG];
    double2* p = tile + stream_chunk2(t, 0);
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) a[r] = p[r << 3];
    if (P.layerA >= 0) {
        double tA[5];
        tile_signed_t(P.t1, P.tb, 3, rmA, tA);
        tile_rot_bits(a, tA, 0, 5);
    }
    if (P.layerD >= 0) {
        const int lane = t & 31, l2 = (lane >> 2) & 1, l8 = (lane >> 3) & 1;
        const double2 cf0 = cmul(tab.T2[t], tab.F[l2][0]), cf1 = cmul(tab.T2[t], tab.F[l2][1]);
        const double2 c00 = cmul(cf0, tab.G[l8][0]), c01 = cmul(cf1, tab.G[l8][0]);
        const double2 c10 = cmul(cf0, tab.G[l8][1]), c11 = cmul(cf1, tab.G[l8][1]);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            a[i] = cmul(a[i], cmul(tab.T1c[i], (i & 1) ? c01 : c00));
            a[i | 16] = cmul(a[i | 16], cmul(tab.T1c[i | 16], (i & 1) ? c11 : c10));
        }
    }
    if (P.layerB >= 0) {
        double tB[5];
        tile_signed_t(P.t2, P.tb, 3, rmB, tB);
        tile_rot_bits(a, tB, 0, 5);
    }
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) p[r << 3] = a[r];
}

// mode C: the only phase.  Thread t <-> passive local bits 0..6, registers <-> local bits 7..11; phase = T1c[r] * T2[t].
template <class PassT>
DTC_HD void stream_phaseC(int t, double2* tile, const StreamSlot& tab, const PassT& P, u64 rmA, u64 rmB) {
    double tA[5], tB[5];
    tile_signed_t(P.t1, P.tb, 7, rmA, tA);
    tile_signed_t(P.t2, P.tb, 7, rmB, tB);
    double2 a[DTC_NREG];
    double2* p = tile + t;
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) a[r] = p[r << 7];
    const double2 c = tab.T2[t];
    if (P.layerA >= 0) tile_rot_bits(a, tA, 0, 5);
    if (P.layerD >= 0) {
#pragma unroll
        for (int r = 0; r < DTC_NREG; ++r) a[r] = cmul(a[r], cmul(tab.T1c[r], c));
    }
    if (P.layerB >= 0) tile_rot_bits(a, tB, 0, 5);
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) p[r << 7] = a[r];
}

// ---- phase tables of one tile, built by ONE warp in three steps separated by __syncwarp().
// g_outer: global index of the tile's local index 0 (incl. rank bits above n_local).
DTC_HD double2 stream_bond(const StreamBuild& bl, const DtcStreamPass& P, int c, int l12) {
    return bl.B[P.Fk[c]][((l12 >> P.Fa[c]) ^ (l12 >> P.Fb[c])) & 1];
}

// step 1: one-body factors E (incl. bonds to outer qubits), bond factors B, outer factors of the tile constant
DTC_HD void stream_build1(int lane, StreamBuild& bl, const DtcStreamPass& P, const DtcLayer& L, u64 g_outer,
                          u64 m1a, u64 m1b, u64 m2) {
    if (lane < 2 * DTC_TILE_BITS) {
        const int l = lane >> 1, bit = lane & 1;
        double2 e = d1_factor(L, P.tb[l], bit, m1a, m1b);
        DTC_NOUNROLL
        for (int c = 0; c < P.nC; ++c)
            if (P.Ca[c] == l) e = cmul(e, d2_factor(L, P.Ck[c], bit ^ (int)((g_outer >> P.Cb[c]) & 1ull), m2));
        bl.E[l][bit] = e;
    }
    DTC_NOUNROLL
    for (int k = lane; k < L.n_terms; k += 32) {
        bl.B[k][0] = d2_factor(L, k, 0, m2);
        bl.B[k][1] = d2_factor(L, k, 1, m2);
    }
    const u64 any = (L.d1_any[0] | L.d1_any[1]) & ~P.tile_mask;
    const int cnt = DTC_POPC64(any);
    double2 f = make_double2(1.0, 0.0);
    DTC_NOUNROLL
    for (int it = lane; it < cnt + P.nO; it += 32) {
        if (it < cnt) {
            u64 m = any;
            DTC_NOUNROLL
            for (int k = 0; k < it; ++k) m &= m - 1;
            const int q = DTC_CTZ64(m);
            f = cmul(f, d1_factor(L, q, (int)((g_outer >> q) & 1ull), m1a, m1b));
        } else {
            const int o = it - cnt;
            const int par = (int)(((g_outer >> P.Oa[o]) ^ (g_outer >> P.Ob[o])) & 1ull);
            f = cmul(f, d2_factor(L, P.Ok[o], par, m2));
        }
    }
    bl.scratch[lane] = f;
}

// step 2: core table (one entry per lane), F/G, and the two halves of T2 (T2lo over local 0..2, T2hi over local 8..11)
DTC_HD void stream_build2(int lane, StreamBuild& bl, const DtcStreamPass& P, const DtcLayer& L) {
    const int cb = stream_core_base(P.mode), hb = stream_hi_base(P.mode);
    {
        const int l12 = lane << cb;
        double2 p = bl.E[cb][lane & 1];
#pragma unroll
        for (int j = 1; j < 5; ++j) p = cmul(p, bl.E[cb + j][(lane >> j) & 1]);
        DTC_NOUNROLL
        for (int c = P.fam_off[0]; c < P.fam_off[1]; ++c) p = cmul(p, stream_bond(bl, P, c, l12));
        bl.T1c[lane] = p;
    }
    if (lane < 16) {
        const int l12 = lane << hb;
        double2 p = bl.E[hb][lane & 1];
#pragma unroll
        for (int j = 1; j < 4; ++j) p = cmul(p, bl.E[hb + j][(lane >> j) & 1]);
        DTC_NOUNROLL
        for (int c = P.fam_off[4]; c < P.fam_off[5]; ++c) p = cmul(p, stream_bond(bl, P, c, l12));
        bl.T2hi[lane] = p;
    } else if (lane < 24) {
        const int i = lane - 16;
        double2 p = cmul(bl.E[0][i & 1], cmul(bl.E[1][(i >> 1) & 1], bl.E[2][(i >> 2) & 1]));
        DTC_NOUNROLL
        for (int c = P.fam_off[3]; c < P.fam_off[4]; ++c) p = cmul(p, stream_bond(bl, P, c, i));
        bl.T2lo[i] = p;
    } else {
        // FG[0..3] = F[l2][l3], FG[4..7] = G[l8][l7]: products of bond factors, index = parity of the two bits
        const int i = lane - 24, fam = 1 + (i >> 2), par = ((i >> 1) ^ i) & 1;
        double2 p = make_double2(1.0, 0.0);
        DTC_NOUNROLL
        for (int c = P.fam_off[fam]; c < P.fam_off[fam + 1]; ++c) p = cmul(p, bl.B[P.Fk[c]][par]);
        bl.FG[i] = p;
        if (lane == 31) {                 // the last lane: the sequential CPU emulation has run all others by now
            const u64 any = (L.d1_any[0] | L.d1_any[1]) & ~P.tile_mask;
            const int n = DTC_POPC64(any) + P.nO;
            double2 c = make_double2(L.cr, L.ci);
            DTC_NOUNROLL
            for (int k = 0; k < (n < 32 ? n : 32); ++k) c = cmul(c, bl.scratch[k]);
            bl.C = c;
        }
    }
}

// step 3: fill the slot
DTC_HD void stream_build3(int lane, const StreamBuild& bl, StreamSlot& slot, const DtcStreamPass& P) {
    const double2 one = make_double2(1.0, 0.0);
    const bool ident = P.layerD < 0;
    slot.T1c[lane] = ident ? one : bl.T1c[lane];
    if (lane < 8) (&slot.F[0][0])[lane] = ident ? one : bl.FG[lane];      // F and G are adjacent: FG[4..7] lands in G
#pragma unroll 1
    for (int m = 0; m < 4; ++m) {
        const int idx = lane + 32 * m;
        if (ident) {
            slot.T2[idx] = one;
            continue;
        }
        const int l12 = (idx & 7) | ((idx >> 3) << stream_hi_base(P.mode));
        double2 q = cmul(bl.C, cmul(bl.T2lo[idx & 7], bl.T2hi[idx >> 3]));
        DTC_NOUNROLL
        for (int c = P.fam_off[5]; c < P.fam_off[6]; ++c) q = cmul(q, stream_bond(bl, P, c, l12));
        slot.T2[idx] = q;
    }
}

// ---- fused read-out (last pass of a factorised circuit): after phase 3 the tile is in the stage buffer; each thread
// takes 16 pairs (l, l | 1 << lb) of the local bit lb and accumulates rho00, rho11, rho01 of psi'.  The stage is not stored.
DTC_HD void stream_rdm_pairs(int t, const double2* tile, int lb, double acc[4]) {
    const int low = (1 << lb) - 1;
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
        const int pi = t + 128 * j;                                  // pair index 0..2047, consecutive lanes adjacent
        const int l0 = ((pi & ~low) << 1) | (pi & low);
        const double2 a = tile[l0], b = tile[l0 | (1 << lb)];
        acc[0] += a.x * a.x + a.y * a.y;
        acc[1] += b.x * b.x + b.y * b.y;
        acc[2] += a.x * b.x + a.y * b.y;                             // a conj(b)
        acc[3] += a.y * b.x - a.x * b.y;
    }
}

struct StreamMasks {
    u64 rmA, rmB, m1a, m1b, m2;
};

DTC_HD StreamMasks stream_load_masks(const DtcStreamPass& P, const u64* masks, long long n_traj, u64 traj) {
    StreamMasks m;
    m.rmA = m.rmB = m.m1a = m.m1b = m.m2 = 0;
    if (P.layerA >= 0) m.rmA = masks[(long long)(P.layerA * 4) * n_traj + traj];
    if (P.layerB >= 0) m.rmB = masks[(long long)(P.layerB * 4) * n_traj + traj];
    if (P.layerD >= 0) {
        m.m1a = masks[(long long)(P.layerD * 4 + 1) * n_traj + traj];
        m.m1b = masks[(long long)(P.layerD * 4 + 2) * n_traj + traj];
        m.m2 = masks[(long long)(P.layerD * 4 + 3) * n_traj + traj];
    }
    return m;
}
