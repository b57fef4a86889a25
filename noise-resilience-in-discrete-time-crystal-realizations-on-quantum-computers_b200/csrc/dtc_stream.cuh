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
// Register sets:  S2 = bits 3..7 in both modes;  S1 = {0,1,2,8,9} (A) / {2,8,9,10,11} (B).
// Bank conflicts: a 16 B access is conflict free when the 8 lanes of a quarter warp hit 8 different
// values of l[0:2].  Phase 2 maps those lanes to l[0:2] directly.  Phases 1/3 hold low bits in
// registers, so lane c reads its registers in the order l_low = k ^ c: the tan-form RX butterfly is
// symmetric under exchanging the two members of a pair, so the permuted order needs no fix-up.
#pragma once
#include "dtc_hd.cuh"

#define DTC_STREAM_STAGES 3
#define DTC_STREAM_WG 2                       // compute warpgroups (128 threads each) per CTA
#define DTC_STREAM_THREADS (128 * DTC_STREAM_WG + 32)

struct DtcStreamPass {
    int mode;                      // 1: A, 2: B
    int contig;                    // tile is 64 KB contiguous in global memory (bulk copy, no tensor map)
    int n_local, g;                // B: tile bits 2..11 = global bits g..g+9
    int layerA, layerD, layerB;
    int n_terms_pad_;
    int tb[DTC_TILE_BITS];
    double t1[DTC_TILE_BITS], t2[DTC_TILE_BITS];
    u64 tile_mask;
    int nT1, nT2, nC, nO;
    unsigned char T1k[DTC_MAXT], T1a[DTC_MAXT], T1b[DTC_MAXT];   // bonds with both ends in local [2, 8]
    unsigned char T2k[DTC_MAXT], T2a[DTC_MAXT], T2b[DTC_MAXT];   // both ends in {0,1,2,8,9,10,11}
    unsigned char Ck[DTC_MAXT], Ca[DTC_MAXT], Cb[DTC_MAXT];      // local a, outer qubit b
    unsigned char Ok[DTC_MAXT], Oa[DTC_MAXT], Ob[DTC_MAXT];      // outer, outer
};

// per-warpgroup phase tables
struct StreamTables {
    double2 T1[128];               // index = local bits 2..8
    double2 T2[128];               // index = local bits 0,1,2,8,9,10,11 (= thread id of phase 2), times the tile constant
    double2 E[DTC_TILE_BITS][2];
    double2 B[DTC_MAXT][2];
    double2 scratch[32];
    double2 C;
};

// ---- tile geometry
DTC_HD u64 stream_tile_base(u64 tile_in_traj, const DtcStreamPass& P) {
    if (P.contig) return tile_in_traj << DTC_TILE_BITS;
    const int lb = P.g - 2;
    return ((tile_in_traj & ((1ull << lb) - 1)) << 2) | ((tile_in_traj >> lb) << (P.g + 10));
}

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

template <int MODE>
DTC_HD void stream_signed_s1(const double* tbase, const int* tb, u64 rmask, double out[5]) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int l = stream_s1_bit<MODE>(j);
        const double t = tbase[l];
        out[j] = ((rmask >> tb[l]) & 1ull) ? -t : t;
    }
}

// ---- phases (tile: the stage buffer, 4096 chunks of 16 B)
template <int MODE>
DTC_HD void stream_phase13(int t, double2* tile, const double* tbase, const int* tb, u64 rmask) {
    double tt[5];
    stream_signed_s1<MODE>(tbase, tb, rmask, tt);
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

DTC_HD void stream_phase2(int t, double2* tile, const StreamTables& tab, const DtcStreamPass& P, u64 rmA, u64 rmB) {
    double tA[5], tB[5];
    tile_signed_t(P.t1, P.tb, 3, rmA, tA);
    tile_signed_t(P.t2, P.tb, 3, rmB, tB);
    double2 a[DTC_NREG];
    double2* p = tile + stream_chunk2(t, 0);
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) a[r] = p[r << 3];
    tile_rot_bits(a, tA, 0, 4);
    const int lane = t & 31;
    const double2 cthr = tab.T2[t];
    // T1 index = l2 | r << 1 | l8 << 6   (l2 = lane bit 2, l8 = lane bit 3)
    const double2* t1 = tab.T1 + (((lane >> 2) & 1) | (((lane >> 3) & 1) << 6));
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        rot_pair(a[i], a[i | 16], tA[4]);
        const double2 p0 = cmul(t1[i << 1], cthr);
        const double2 p1 = cmul(t1[(i | 16) << 1], cthr);
        a[i] = cmul(a[i], p0);
        a[i | 16] = cmul(a[i | 16], p1);
        rot_pair(a[i], a[i | 16], tB[4]);
        DTC_SCHED_FENCE();
    }
    tile_rot_bits(a, tB, 0, 4);
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) p[r << 3] = a[r];
}

// ---- phase tables of one tile.  Step 1 (E, B, tile constant), warpgroup barrier, step 2 (T1, T2).
// g_outer: global index of the tile's local index 0 (incl. rank bits above n_local).
DTC_HD void stream_setup1(int t, StreamTables& tab, const DtcStreamPass& P, const DtcLayer& L, u64 g_outer,
                          u64 m1a, u64 m1b, u64 m2) {
    if (t < 2 * DTC_TILE_BITS) {
        const int l = t >> 1, bit = t & 1;
        double2 e = d1_factor(L, P.tb[l], bit, m1a, m1b);
        for (int c = 0; c < P.nC; ++c)
            if (P.Ca[c] == l) e = cmul(e, d2_factor(L, P.Ck[c], bit ^ (int)((g_outer >> P.Cb[c]) & 1ull), m2));
        tab.E[l][bit] = e;
    } else if (t >= 32 && t < 32 + DTC_MAXT) {
        const int k = t - 32;
        if (k < L.n_terms) {
            tab.B[k][0] = d2_factor(L, k, 0, m2);
            tab.B[k][1] = d2_factor(L, k, 1, m2);
        }
    } else if (t >= 96) {
        const int lane = t - 96;
        const u64 any = (L.d1_any[0] | L.d1_any[1]) & ~P.tile_mask;
        const int cnt = DTC_POPC64(any);
        double2 f = make_double2(1.0, 0.0);
        for (int it = lane; it < cnt + P.nO; it += 32) {
            if (it < cnt) {
                u64 m = any;
                for (int k = 0; k < it; ++k) m &= m - 1;
                const int q = DTC_CTZ64(m);
                f = cmul(f, d1_factor(L, q, (int)((g_outer >> q) & 1ull), m1a, m1b));
            } else {
                const int o = it - cnt;
                const int par = (int)(((g_outer >> P.Oa[o]) ^ (g_outer >> P.Ob[o])) & 1ull);
                f = cmul(f, d2_factor(L, P.Ok[o], par, m2));
            }
        }
        tab.scratch[lane] = f;
        DTC_SYNCWARP();
        if (lane == 31) {                 // the last lane: the sequential CPU emulation has run all others by now
            double2 c = make_double2(L.cr, L.ci);
            const int used = (cnt + P.nO < 32) ? cnt + P.nO : 32;
            for (int k = 0; k < used; ++k) c = cmul(c, tab.scratch[k]);
            tab.C = c;
        }
    }
}

DTC_HD void stream_setup2(int t, StreamTables& tab, const DtcStreamPass& P) {
    if (P.layerD < 0) {
        tab.T1[t] = make_double2(1.0, 0.0);
        tab.T2[t] = make_double2(1.0, 0.0);
        return;
    }
    {   // T1: index bit j <-> local bit 2 + j; one-body factors of local bits 3..7, bonds inside [2, 8]
        double2 p = tab.E[3][(t >> 1) & 1];
#pragma unroll
        for (int l = 4; l <= 7; ++l) p = cmul(p, tab.E[l][(t >> (l - 2)) & 1]);
        for (int c = 0; c < P.nT1; ++c)
            p = cmul(p, tab.B[P.T1k[c]][((t >> (P.T1a[c] - 2)) ^ (t >> (P.T1b[c] - 2))) & 1]);
        tab.T1[t] = p;
    }
    {   // T2: index bits 0..2 <-> local 0..2, bits 3..6 <-> local 8..11
        const int l12 = (t & 7) | ((t >> 3) << 8);          // the thread's fixed local bits as a 12-bit index
        double2 p = tab.C;
#pragma unroll
        for (int l = 0; l < DTC_TILE_BITS; ++l)
            if (l < 3 || l > 7) p = cmul(p, tab.E[l][(l12 >> l) & 1]);
        for (int c = 0; c < P.nT2; ++c)
            p = cmul(p, tab.B[P.T2k[c]][((l12 >> P.T2a[c]) ^ (l12 >> P.T2b[c])) & 1]);
        tab.T2[t] = p;
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
