// dtc_hd.cuh -- host/device shared building blocks of the dtcsim kernels.
//
// Everything here compiles both under nvcc (device code of dtcsim.cu) and under plain g++ (the
// CPU emulation harness tests/emul/emul.cpp that executes the *same* per-thread code thread by
// thread, phase by phase, so the kernel logic is testable without a GPU).
//
// Hot path this implements (reference: the per-shot statevector evolution inside
// AerSimulator.run(), fast.py:211; circuit shape fast.py:111-147):
//   * RX layer  -> tan-form butterflies  out0 = x0 - i t x1, out1 = x1 - i t x0  (4 DFMA / pair; the
//                  cos factors are folded into the next diagonal layer's constant)
//   * ZZ(phis) / Z(hs) diagonal layer -> product of two shared-memory phase tables
//   * depolarizing noise -> Philox-sampled Pauli frames that only flip angle signs
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DTC_HD __host__ __device__ __forceinline__
#else
#define DTC_HD inline
struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) {
    double2 r;
    r.x = x;
    r.y = y;
    return r;
}
#endif

typedef unsigned long long u64;

#define DTC_MAXQ 64
#define DTC_MAXT 64
#define DTC_TILE_BITS 12
#define DTC_TILE (1 << DTC_TILE_BITS)
#define DTC_NREG 32
#define DTC_THREADS 128

enum { DTC_EVT_ROT = 0, DTC_EVT_D1 = 1, DTC_EVT_D2 = 2, DTC_EVT_NOISE = 3, DTC_EVT_D2C = 4 };
#define DTC_VIRTUAL_QUBIT 63   // table partner of D2C terms (partner qubit still |0> in psi'): index bit always 0

// Per-layer tables (R_j followed by D_j), device resident, shared by all trajectories.
struct DtcLayer {
    double cr, ci;                 // layer constant: prod cos(theta'/2) (x e^{i global phase} in layer 0)
    double c1[2][DTC_MAXQ];        // cos(a/2) of D1 slot s on qubit q (1 if absent)
    double s1[2][DTC_MAXQ];        // sin(a/2)                         (0 if absent)
    double tc[DTC_MAXT], ts[DTC_MAXT];   // cos(b/2), sin(b/2) of D2 term k
    int ti[DTC_MAXT], tj[DTC_MAXT];      // its qubits
    u64 d1_any[2];                 // qubits with a non-trivial slot-s coefficient
    int n_terms;
    int pad_;
    // ---- everything above is what the device-side table builders read (DTC_LAYER_PREFIX bytes); below: host / frames only
    u64 rot_any;                   // qubits rotated in R_j
    double rtan[DTC_MAXQ];         // tan(theta'/2) of the rotation on q in R_j (0 if none)
};
#define DTC_LAYER_PREFIX (16 + 4 * 8 * DTC_MAXQ + 2 * 8 * DTC_MAXT + 2 * 4 * DTC_MAXT + 16 + 8)

struct DtcEvent {
    int type, layer, q0, q1, slot, k;    // k: ROT quarter turns (theta = theta' + k pi)
    double c0, c1, c2;                   // NOISE cumulative probabilities pX, pX+pY, pX+pY+pZ
};

// One fused pass  R_A|S -> D -> R_B|S  over tiles of 2^12 amplitudes.
struct DtcTilePass {
    int n_local, n_total;
    int s2_lo;                     // layout: S2 = local bits [s2_lo, s2_lo+5), S1 = [s2_lo+5, s2_lo+10)
    int layerA, layerD, layerB;    // layer ids (-1: absent)
    int tb[DTC_TILE_BITS];         // global bit position of tile-local bit l (ascending)
    double t1[DTC_TILE_BITS];      // tan for R_A on local bit l (0: none in this pass)
    double t2[DTC_TILE_BITS];      // tan for R_B
    u64 roff[DTC_NREG];            // global offset of register r in phases 1/3 (sum of S1 strides)
    u64 tid_off[7];                // global offset contributed by thread-id bit b in phases 1/3
    u64 tile_mask;                 // global bit positions covered by the tile
    int seg_n;                     // tile counter -> global base: sum over segments of
    int seg_src[8], seg_len[8], seg_dst[8];   //   ((tile >> src) & ((1 << len) - 1)) << dst
    // classification of D_layerD's two-body terms relative to this tile
    int nT1, nT2, nX, nC, nO;
    unsigned char T1k[DTC_MAXT], T1a[DTC_MAXT], T1b[DTC_MAXT];   // both ends in local [0, s1_lo]
    unsigned char T2k[DTC_MAXT], T2a[DTC_MAXT], T2b[DTC_MAXT];   // both ends in local [s1_lo, 12)
    unsigned char Xk[DTC_MAXT], Xa[DTC_MAXT], Xb[DTC_MAXT];      // local-local, not in one table
    unsigned char Ck[DTC_MAXT], Ca[DTC_MAXT], Cb[DTC_MAXT];      // local a, outer qubit b
    unsigned char Ok[DTC_MAXT], Oa[DTC_MAXT], Ob[DTC_MAXT];      // outer, outer
};

DTC_HD double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// ------------------------------------------------------------------------------------ Philox
DTC_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const u64 p0 = (u64)0xD2511F53u * c0;
        const u64 p1 = (u64)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// stream 0: noise site `index`; stream 1: measurement sample `index` (contract in oracle/philox.py)
DTC_HD double philox_uniform(u64 seed, uint32_t index, uint32_t stream, u64 traj) {
    uint32_t o[4];
    philox4x32_10(index, stream, (uint32_t)traj, (uint32_t)(traj >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const u64 bits = (u64)o[0] | ((u64)o[1] << 32);
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------ frames
// Walk the event list for one trajectory; masks[(layer*4+m)*mstride] |= bits; returns final frame.
// m = 0: rotation sign bits (bit q), 1/2: D1 slot 0/1 (bit q), 3: D2 (bit = term index).
// one event of the walk; the frame (fx, fz, ph) is carried by the caller
DTC_HD void frame_step(const DtcEvent& E, u64 traj_global, u64 seed, u64* masks, int64_t mstride, u64& fx, u64& fz, int& ph) {
    const int q = E.q0;
    if (E.type == DTC_EVT_ROT) {
        if (E.slot == 0 && ((fz >> q) & 1ull)) masks[(int64_t)(E.layer * 4 + 0) * mstride] |= 1ull << q;
        if (E.k & 1) fx ^= 1ull << q;
        ph += 3 * E.k;
    } else if (E.type == DTC_EVT_D1) {
        if ((fx >> q) & 1ull) masks[(int64_t)(E.layer * 4 + 1 + E.slot) * mstride] |= 1ull << q;
    } else if (E.type == DTC_EVT_D2 || E.type == DTC_EVT_D2C) {
        if (((fx >> q) ^ (fx >> E.q1)) & 1ull) masks[(int64_t)(E.layer * 4 + 3) * mstride] |= 1ull << E.slot;
    } else {
        const double u = philox_uniform(seed, (uint32_t)E.slot, 0u, traj_global);
        const int fxq = (int)((fx >> q) & 1ull);
        if (u < E.c0) {                    // X
            fx ^= 1ull << q;
        } else if (u < E.c1) {             // Y = i X Z
            ph += 1 + 2 * fxq;
            fx ^= 1ull << q;
            fz ^= 1ull << q;
        } else if (u < E.c2) {             // Z
            ph += 2 * fxq;
            fz ^= 1ull << q;
        }
    }
}

DTC_HD void frame_walk(u64 traj_global, u64 seed, const DtcEvent* ev, int64_t n_events, u64* masks,
                       int64_t mstride, u64* fx_out, u64* fz_out, int* ph_out) {
    u64 fx = 0, fz = 0;
    int ph = 0;
    for (int64_t e = 0; e < n_events; ++e) frame_step(ev[e], traj_global, seed, masks, mstride, fx, fz, ph);
    *fx_out = fx;
    *fz_out = fz;
    *ph_out = ph & 3;
}

// ------------------------------------------------------------------------------------ generic engine
// sign-resolved D1 factor of qubit q for bit value `bit` (z = 1 - 2 bit)
DTC_HD double2 d1_factor(const DtcLayer& L, int q, int bit, u64 m1a, u64 m1b) {
    const double z = bit ? -1.0 : 1.0;
    const double sa = ((m1a >> q) & 1ull) ? -z : z;
    const double sb = ((m1b >> q) & 1ull) ? -z : z;
    const double2 fa = make_double2(L.c1[0][q], -sa * L.s1[0][q]);
    const double2 fb = make_double2(L.c1[1][q], -sb * L.s1[1][q]);
    return cmul(fa, fb);
}

DTC_HD double2 d2_factor(const DtcLayer& L, int k, int parity, u64 m2) {
    const double z = parity ? -1.0 : 1.0;
    const double s = ((m2 >> k) & 1ull) ? -z : z;
    return make_double2(L.tc[k], -s * L.ts[k]);
}

// full phase of D_layer at global basis index g (generic engine and per-tile constants)
DTC_HD double2 diag_phase(const DtcLayer& L, u64 g, u64 m1a, u64 m1b, u64 m2) {
    double2 p = make_double2(L.cr, L.ci);
    u64 any = L.d1_any[0] | L.d1_any[1];
    while (any) {
#if defined(__CUDA_ARCH__)
        const int q = __ffsll((long long)any) - 1;
#else
        const int q = __builtin_ctzll(any);
#endif
        any &= any - 1;
        p = cmul(p, d1_factor(L, q, (int)((g >> q) & 1ull), m1a, m1b));
    }
    for (int k = 0; k < L.n_terms; ++k) {
        const int par = (int)(((g >> L.ti[k]) ^ (g >> L.tj[k])) & 1ull);
        p = cmul(p, d2_factor(L, k, par, m2));
    }
    return p;
}

// tan-form RX butterfly on a pair
DTC_HD void rot_pair(double2& x0, double2& x1, double t) {
    const double2 a = x0, b = x1;
    x0.x = fma(t, b.y, a.x);
    x0.y = fma(-t, b.x, a.y);
    x1.x = fma(t, a.y, b.x);
    x1.y = fma(-t, a.x, b.y);
}

// ------------------------------------------------------------------------------------ tile engine
#define DTC_TILE_PAD (DTC_TILE + 128)     // phys(l) = l + (l >> S1), S1 >= 5
struct TileSmem {
    double2 tile[DTC_TILE_PAD];
    double2 T1[256 + 8];
    double2 T2[128];
    double2 E[DTC_TILE_BITS][2];
    double2 B[DTC_MAXT][2];
    double2 C;
    u64 base;          // global index of tile-local index 0 (kept here, not in registers, across phases)
    u64 rmA, rmB;      // rotation sign masks of the trajectory
};

// Shared-memory layout: one padding chunk (16 B) after every 2^S1 chunks.  For l = T | R (thread bits and
// register bits disjoint) phys(l) = phys(T) + phys(R), so every access is thread base + compile-time constant;
// quarter-warps hit 8 distinct 16 B bank groups in all three phases (no conflicts).
template <int S2_LO>
DTC_HD int tile_pad(int l) { return l + (l >> (S2_LO + 5)); }

// local index of (thread, register) in phases 1 and 3 (register bits = S1)
template <int S2_LO>
DTC_HD int tile_local_p13(int tid, int r) {
    constexpr int S1 = S2_LO + 5;
    return (tid & ((1 << S1) - 1)) | (r << S1) | ((tid >> S1) << (S1 + 5));
}
// ... in phase 2 (register bits = S2): low thread bits <-> local bits >= S1, high thread bits <-> local bits < S2_LO
template <int S2_LO>
DTC_HD int tile_local_p2(int tid, int r) {
    constexpr int S1 = S2_LO + 5;
    constexpr int NH = 7 - S2_LO;
    return (r << S2_LO) | ((tid & ((1 << NH) - 1)) << S1) | (tid >> NH);
}

DTC_HD u64 tile_base_index(u64 tile, const DtcTilePass& P) {
    u64 base = 0;
    for (int k = 0; k < P.seg_n; ++k)
        base |= ((tile >> P.seg_src[k]) & ((1ull << P.seg_len[k]) - 1)) << P.seg_dst[k];
    return base;
}

// rotations on the five register bits; t[k] already carries the trajectory's sign
DTC_HD void tile_rot5(double2 a[DTC_NREG], const double t[5]) {
    // no "skip if t == 0" branches: x0 - i*0*x1 is exact, and straight-line code lets ptxas rename
    // registers instead of moving every result back to a home register
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int i = 0; i < DTC_NREG; ++i) {
            if (!((i >> k) & 1)) rot_pair(a[i], a[i | (1 << k)], t[k]);
        }
    }
}

#if defined(__CUDA_ARCH__)
#define DTC_SYNCWARP() __syncwarp()
#define DTC_CTZ64(x) (__ffsll((long long)(x)) - 1)
#define DTC_POPC64(x) __popcll(x)
#else
#define DTC_SYNCWARP() ((void)0)
#define DTC_CTZ64(x) __builtin_ctzll(x)
#define DTC_POPC64(x) __builtin_popcountll(x)
#endif

// per-CTA setup of the diagonal layer: E (local one-body incl. cross terms), B (local two-body), C.
// Warp 0 lanes 0..23: E; warps 1,2: B; warp 3: tile constant C, one factor per lane (outer one-body
// factors and outer-outer bonds), multiplied together by the warp's last lane.
// g_outer: global index with all tile-local bits 0.
DTC_HD void tile_setup_thread(int tid, TileSmem& sm, const DtcTilePass& P, const DtcLayer& L, u64 g_outer,
                              u64 m1a, u64 m1b, u64 m2) {
    if (tid < 2 * DTC_TILE_BITS) {
        const int l = tid >> 1, bit = tid & 1;
        const int q = P.tb[l];
        double2 e = d1_factor(L, q, bit, m1a, m1b);
        for (int c = 0; c < P.nC; ++c) {
            if (P.Ca[c] == l) {
                const int par = bit ^ (int)((g_outer >> P.Cb[c]) & 1ull);
                e = cmul(e, d2_factor(L, P.Ck[c], par, m2));
            }
        }
        sm.E[l][bit] = e;
    } else if (tid >= 32 && tid < 32 + DTC_MAXT) {
        const int k = tid - 32;
        if (k < L.n_terms) {
            sm.B[k][0] = d2_factor(L, k, 0, m2);
            sm.B[k][1] = d2_factor(L, k, 1, m2);
        }
    } else if (tid >= 96) {
        const int lane = tid - 96;
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
        sm.T2[lane] = f;                    // scratch: T2 is rebuilt after the next barrier
        DTC_SYNCWARP();
        if (lane == 31) {
            double2 c = make_double2(L.cr, L.ci);
            const int used = (cnt + P.nO < 32) ? cnt + P.nO : 32;
            for (int k = 0; k < used; ++k) c = cmul(c, sm.T2[k]);
            sm.C = c;
        }
    }
}

// phase tables: T1 over local bits [0, S1_LO] (one-body of all of them), T2 over [S1_LO, 12)
// (one-body of (S1_LO, 12) only; bit S1_LO enters T2 through two-body terms).
template <int S2_LO>
DTC_HD void tile_tables_thread(int tid, TileSmem& sm, const DtcTilePass& P) {
    constexpr int S1 = S2_LO + 5;
    constexpr int T1B = S1 + 1;
    constexpr int T2B = DTC_TILE_BITS - S1;
    if (P.layerD < 0) {                       // pass without a diagonal layer: identity tables
        for (int idx = tid; idx < (1 << T1B); idx += DTC_THREADS) sm.T1[idx + 4 * (idx >> S1)] = make_double2(1.0, 0.0);
        for (int idx = tid; idx < (1 << T2B); idx += DTC_THREADS) sm.T2[idx] = make_double2(1.0, 0.0);
        if (tid == 0) sm.C = make_double2(1.0, 0.0);
        return;
    }
    for (int idx = tid; idx < (1 << T1B); idx += DTC_THREADS) {
        double2 p = sm.E[0][idx & 1];
#pragma unroll
        for (int l = 1; l < T1B; ++l) p = cmul(p, sm.E[l][(idx >> l) & 1]);
        for (int c = 0; c < P.nT1; ++c)
            p = cmul(p, sm.B[P.T1k[c]][((idx >> P.T1a[c]) ^ (idx >> P.T1b[c])) & 1]);
        sm.T1[idx + 4 * (idx >> S1)] = p;     // +4 chunks when bit S1 is set: conflict-free lookups
    }
    for (int idx = tid; idx < (1 << T2B); idx += DTC_THREADS) {
        double2 p = make_double2(1.0, 0.0);
#pragma unroll
        for (int m = 1; m < T2B; ++m) p = cmul(p, sm.E[S1 + m][(idx >> m) & 1]);
        for (int c = 0; c < P.nT2; ++c)
            p = cmul(p, sm.B[P.T2k[c]][((idx >> (P.T2a[c] - S1)) ^ (idx >> (P.T2b[c] - S1))) & 1]);
        sm.T2[idx] = p;
    }
}

// resolve the trajectory's signs for the five register bits starting at local bit lo
template <class TB>
DTC_HD void tile_signed_t(const double* tbase, const TB* tb, int lo, u64 rmask, double out[5]) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double t = tbase[lo + k];
        out[k] = ((rmask >> tb[lo + k]) & 1ull) ? -t : t;
    }
}

#ifndef DTC_FENCE_EVERY
#define DTC_FENCE_EVERY 1      // pairs between scheduling fences in the fused diagonal loop (0: none)
#endif
#if defined(__CUDA_ARCH__)
#define DTC_SCHED_FENCE() asm volatile("" ::: "memory")
#else
#define DTC_SCHED_FENCE() ((void)0)
#endif

// rotations on register bits [kbeg, kend)
DTC_HD void tile_rot_bits(double2 a[DTC_NREG], const double t[5], int kbeg, int kend) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (k < kbeg || k >= kend) continue;
#pragma unroll
        for (int i = 0; i < DTC_NREG; ++i) {
            if (!((i >> k) & 1)) rot_pair(a[i], a[i | (1 << k)], t[k]);
        }
    }
}

// phase 2 body on the register file: R_A|S2, D, R_B|S2 -- branch-free.  Absent layers have t = 0 /
// identity tables.  The diagonal multiply is fused into the pair loop of the top register bit
// (R_A bit 4 -> D -> R_B bit 4 per pair) so that only two phase-table entries are live at a time.
template <int S2_LO, bool HAS_X>
DTC_HD void tile_phase2_compute(int tid, double2 a[DTC_NREG], const TileSmem& sm, const DtcTilePass& P,
                                u64 rmA, u64 rmB) {
    constexpr int S1 = S2_LO + 5;
    constexpr int NH = 7 - S2_LO;
    double tA[5], tB[5];
    tile_signed_t(P.t1, P.tb, S2_LO, rmA, tA);
    tile_signed_t(P.t2, P.tb, S2_LO, rmB, tB);
    tile_rot_bits(a, tA, 0, 4);
    const double2 cthr = cmul(sm.C, sm.T2[tid & ((1 << NH) - 1)]);
    // T1 index = passive low bits | register bits << S2_LO | bit S1 (thread bit 0), padded by 4*(bit S1)
    const double2* t1 = sm.T1 + ((tid >> NH) | ((tid & 1) << S1)) + 4 * (tid & 1);
    if (!HAS_X) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            rot_pair(a[i], a[i | 16], tA[4]);
            const double2 p0 = cmul(t1[i << S2_LO], cthr);
            const double2 p1 = cmul(t1[(i | 16) << S2_LO], cthr);
            a[i] = cmul(a[i], p0);
            a[i | 16] = cmul(a[i | 16], p1);
            rot_pair(a[i], a[i | 16], tB[4]);
            if (DTC_FENCE_EVERY > 0 && (i % (DTC_FENCE_EVERY > 0 ? DTC_FENCE_EVERY : 1)) == DTC_FENCE_EVERY - 1) DTC_SCHED_FENCE();
        }
    } else {
        tile_rot_bits(a, tA, 4, 5);
#pragma unroll
        for (int r = 0; r < DTC_NREG; ++r) a[r] = cmul(a[r], cmul(t1[r << S2_LO], cthr));
        for (int c = 0; c < P.nX; ++c) {                          // terms outside both tables (rare)
            const double2 b0 = sm.B[P.Xk[c]][0], b1 = sm.B[P.Xk[c]][1];
#pragma unroll
            for (int r = 0; r < DTC_NREG; ++r) {
                const int l = tile_local_p2<S2_LO>(tid, r);
                a[r] = cmul(a[r], (((l >> P.Xa[c]) ^ (l >> P.Xb[c])) & 1) ? b1 : b0);
            }
        }
        tile_rot_bits(a, tB, 4, 5);
    }
    tile_rot_bits(a, tB, 0, 4);
}

// ---- data movement of the three phases (shared with the CPU emulation harness)
#if defined(__CUDA_ARCH__)
#define DTC_LDG(p) __ldcg(p)
#define DTC_STG(p, v) __stcg(p, v)
#else
#define DTC_LDG(p) (*(p))
#define DTC_STG(p, v) (*(p) = (v))
#endif

DTC_HD u64 tile_thread_offset(int tid, u64 base, const DtcTilePass& P) {
    u64 o = base;
#pragma unroll
    for (int b = 0; b < 7; ++b)
        if ((tid >> b) & 1) o += P.tid_off[b];
    return o;
}

// register offsets come from the pass descriptor (constant bank), so no address registers stay live
DTC_HD void tile_gload(const double2* st, u64 off, const DtcTilePass& P, double2 a[DTC_NREG]) {
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) a[r] = DTC_LDG(st + (off + P.roff[r]));
}

DTC_HD void tile_gstore(double2* st, u64 off, const DtcTilePass& P, const double2 a[DTC_NREG]) {
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) DTC_STG(st + (off + P.roff[r]), a[r]);
}

template <int S2_LO>
DTC_HD void tile_sm_store13(int tid, TileSmem& sm, const double2 a[DTC_NREG]) {
    double2* p = sm.tile + tile_pad<S2_LO>(tile_local_p13<S2_LO>(tid, 0));
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) p[tile_pad<S2_LO>(r << (S2_LO + 5))] = a[r];
}
template <int S2_LO>
DTC_HD void tile_sm_load13(int tid, const TileSmem& sm, double2 a[DTC_NREG]) {
    const double2* p = sm.tile + tile_pad<S2_LO>(tile_local_p13<S2_LO>(tid, 0));
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) a[r] = p[tile_pad<S2_LO>(r << (S2_LO + 5))];
}
template <int S2_LO>
DTC_HD void tile_sm_store2(int tid, TileSmem& sm, const double2 a[DTC_NREG]) {
    double2* p = sm.tile + tile_pad<S2_LO>(tile_local_p2<S2_LO>(tid, 0));
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) p[r << S2_LO] = a[r];
}
template <int S2_LO>
DTC_HD void tile_sm_load2(int tid, const TileSmem& sm, double2 a[DTC_NREG]) {
    const double2* p = sm.tile + tile_pad<S2_LO>(tile_local_p2<S2_LO>(tid, 0));
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) a[r] = p[r << S2_LO];
}

// rotations of layer tables tbase on the S1 register bits
template <int S2_LO>
DTC_HD void tile_rot_s1(double2 a[DTC_NREG], const double* tbase, const int* tb, u64 rmask) {
    double t[5];
    tile_signed_t(tbase, tb, S2_LO + 5, rmask, t);
    tile_rot5(a, t);
}

// 228 KB of shared memory per SM, 1 KB reserved per CTA: three resident CTAs need <= 75 KB each
static_assert(3 * (sizeof(TileSmem) + 1024) <= 228 * 1024, "three CTAs per SM must fit in shared memory");

struct TileMasks {
    u64 rmA, rmB, m1a, m1b, m2;
};

DTC_HD TileMasks tile_load_masks(const DtcTilePass& P, const u64* masks, long long n_traj, u64 traj) {
    TileMasks m;
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

// L2 prefetch of the tile a later CTA will load (one request per 64 B run)
#if defined(__CUDA_ARCH__)
#define DTC_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define DTC_PREFETCH_L2(p) ((void)(p))
#endif

DTC_HD void tile_prefetch(const double2* st, u64 off, const DtcTilePass& P) {
#pragma unroll
    for (int r = 0; r < DTC_NREG; ++r) DTC_PREFETCH_L2(st + (off + P.roff[r]));
}
