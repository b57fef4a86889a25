// dtc_dm.cuh -- exact density-matrix evolution (Aer method density_matrix: what AerSimulator.run() does for the reference's
// L = 4 configuration, fast.py:211 with shots > 2^n): pass planner + per-thread code of the fused sweeps of rho.
//
// Compiles under nvcc (kernels k_dm_tile / k_dm_reg of dtcsim.cu) and under g++ (tests/emul/emul.cpp executes the same planner and
// the same per-thread functions CTA by CTA, round by round).
//
// rho is a 2n-bit vector (index = row + 2^n col).  One sweep applies, to up to six qubits, RX(theta) on the row bit, its
// conjugate on the column bit and the Pauli channel that follows; the diagonal layer before it enters at load time through a
// 2^n-entry table T[x] = exp(-i phi(x)):  rho[r,c] *= T[r] conj(T[c]).
//
// k_dm_reg (round 3 of the kernel's history; tiles of 2^12 / 2^13 elements): a thread keeps 16 elements = the (row, column) bits
// of TWO qubits in registers.  Round 0 loads them straight from HBM (phase table applied on the way), rotates, mixes and writes
// the tile to shared memory; the middle round goes shared -> registers -> shared; the last round goes shared -> registers -> HBM.
// Per tile: 4 shared-memory transfers and 2 barriers (the element-per-thread version k_dm_tile: 14 and 8).  Rotations are in tan
// form (out0 = x0 -+ i t x1; |t| <= 1, a quarter turn becomes a swap folded into the channel constants) with cos^2 folded into
// the channel: 4 + 2..4 FP64 operations per element and qubit instead of 12.
#pragma once
#include <string>
#include <vector>

#include "dtc_hd.cuh"

#define DTC_DM_MAXQ 6
struct DmQubitOp {
    int lr, lc;                      // tile-local positions of the qubit's row / column bit (lr < lc)
    double c, s;                     // cos, sin of theta/2 (row: RX(theta), column: its conjugate)
    double dA, dB, oA, oB;           // channel: diagonal block mixing (r == c), off-diagonal block mixing (r != c)
};
struct DmTilePass {
    int n, tile_bits, nq, has_diag;
    int tb[13];                      // global bit of tile-local bit l
    int seg_n, seg_src[8], seg_len[8], seg_dst[8];     // CTA index -> global base (bits outside the tile)
    DmQubitOp q[DTC_DM_MAXQ];
};
struct DmDiagTerms {
    int n1, n2;
    int q1[16];
    double a[16];
    int qi[DTC_MAXT], qj[DTC_MAXT];
    double b[DTC_MAXT];
};

// ---- general single-qubit channel (non-Pauli part of a device noise model): 4 x 4 complex superoperator on the block
// (row bit q, column bit q) with block index = row + 2 col; thread i handles the i-th block
struct DmSuperop { double re[16], im[16]; };
DTC_HD void dm_superop_thread(double2* rho, int n, int q, const DmSuperop& S, long long i) {
    const int b0 = q, b1 = q + n;
    u64 x = (u64)i;
    const u64 low0 = (1ull << b0) - 1;
    x = ((x & ~low0) << 1) | (x & low0);
    const u64 low1 = (1ull << b1) - 1;
    x = ((x & ~low1) << 1) | (x & low1);
    const u64 idx[4] = {x, x | (1ull << b0), x | (1ull << b1), x | (1ull << b0) | (1ull << b1)};
    double2 e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) e[k] = rho[idx[k]];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double sr = S.re[4 * r + c], si = S.im[4 * r + c];
            acc.x = fma(sr, e[c].x, fma(-si, e[c].y, acc.x));
            acc.y = fma(sr, e[c].y, fma(si, e[c].x, acc.y));
        }
        rho[idx[r]] = acc;
    }
}

// ---- register-resident passes
struct DmRegOp {
    double t;                        // tan form of the rotation (0: none)
    double A, B, C, D;               // out(r==c) = A e + B e_flipped ; out(r!=c) = C e + D e_flipped  (cos^2 / swap folded in)
};
struct DmRegRound {
    int hb[4];                       // tile-local bits held in registers: register index bit 0 = row of qubit a, 1 = its column,
                                     // 2 = row of qubit b, 3 = its column (spare bits ride along when nop == 1)
    int tbit[9];                     // tile-local bit carried by thread-index bit k
    int nop, pad_;
    // derived by dm_finish_round (what the kernel reads): tile-local / global element offset of thread-index bit k and of the
    // four register-index bits (global offsets fit 32 bits: n <= 13)
    int emask[9], hp[4];             // hp: swizzled shared-memory offset (dm_phys of the held bit)
    unsigned gmask[9], hg[4];
    DmRegOp op[2];
};
struct DmRegPass {
    int n, tile_bits, n_rounds, has_diag;
    int tb[13];
    int seg_n, seg_src[8], seg_len[8], seg_dst[8];
    DmRegRound r[3];
};

// shared-memory position of tile element e: XOR swizzle of the low three bits with bits 3..5 (16 B elements: a quarter warp
// is conflict free when its eight lanes differ in phys[0:3])
DTC_HD int dm_phys(int e) { return e ^ ((e >> 3) & 7); }

DTC_HD u64 dm_cta_base(u64 cta, int seg_n, const int* seg_src, const int* seg_len, const int* seg_dst) {
    u64 base = 0;
    for (int k = 0; k < seg_n; ++k) base |= ((cta >> seg_src[k]) & ((1ull << seg_len[k]) - 1)) << seg_dst[k];
    return base;
}

#if defined(__CUDA_ARCH__)
#define DMR_LOAD_STREAM(p) __ldcs(p)
#define DMR_STORE_STREAM(p, v) __stcs(p, v)
#define DMR_LOAD_TABLE(p) __ldg(p)
#else
#define DMR_LOAD_STREAM(p) (*(p))
#define DMR_STORE_STREAM(p, v) (*(p) = (v))
#define DMR_LOAD_TABLE(p) (*(p))
#endif

// rotation + conjugate + channel of one qubit on a thread's 16 elements; JR / JC = register-index bits of its row / column bit
template <int JR, int JC>
DTC_HD void dmr_apply_qubit(double2* v, const DmRegOp& o) {
    constexpr int mr = 1 << JR, mc = 1 << JC;
    const double t = o.t;
    const bool rot = t != 0.0, mixd = o.B != 0.0, mixo = o.D != 0.0;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        if (g & (mr | mc)) continue;
        double2 e00 = v[g], e10 = v[g | mr], e01 = v[g | mc], e11 = v[g | mr | mc];      // e<row bit><column bit>
        if (rot) {
            // rows: x0 - i t x1, x1 - i t x0 for each column bit
            const double2 a00 = make_double2(fma(t, e10.y, e00.x), fma(-t, e10.x, e00.y));
            const double2 a10 = make_double2(fma(t, e00.y, e10.x), fma(-t, e00.x, e10.y));
            const double2 a01 = make_double2(fma(t, e11.y, e01.x), fma(-t, e11.x, e01.y));
            const double2 a11 = make_double2(fma(t, e01.y, e11.x), fma(-t, e01.x, e11.y));
            // columns: the conjugate, x0 + i t x1, for each row bit
            e00 = make_double2(fma(-t, a01.y, a00.x), fma(t, a01.x, a00.y));
            e01 = make_double2(fma(-t, a00.y, a01.x), fma(t, a00.x, a01.y));
            e10 = make_double2(fma(-t, a11.y, a10.x), fma(t, a11.x, a10.y));
            e11 = make_double2(fma(-t, a10.y, a11.x), fma(t, a10.x, a11.y));
        }
        if (mixd) {
            v[g] = make_double2(fma(o.B, e11.x, o.A * e00.x), fma(o.B, e11.y, o.A * e00.y));
            v[g | mr | mc] = make_double2(fma(o.B, e00.x, o.A * e11.x), fma(o.B, e00.y, o.A * e11.y));
        } else {
            v[g] = make_double2(o.A * e00.x, o.A * e00.y);
            v[g | mr | mc] = make_double2(o.A * e11.x, o.A * e11.y);
        }
        if (mixo) {
            v[g | mr] = make_double2(fma(o.D, e01.x, o.C * e10.x), fma(o.D, e01.y, o.C * e10.y));
            v[g | mc] = make_double2(fma(o.D, e10.x, o.C * e01.x), fma(o.D, e10.y, o.C * e01.y));
        } else {
            v[g | mr] = make_double2(o.C * e10.x, o.C * e10.y);
            v[g | mc] = make_double2(o.C * e01.x, o.C * e01.y);
        }
    }
}

// One round of a register pass for thread `tid` of the CTA whose tile starts at element `base`: load (HBM in round 0, shared
// memory later), the round's one or two qubits, store (shared memory, HBM in the last round).  The caller puts a barrier between
// rounds.  LB = thread-index bits = tile_bits - 4.  Branch-free index arithmetic on 32-bit element offsets.
template <int LB>
DTC_HD void dmr_round(int r, int tid, const DmRegPass& P, double2* rho, const double2* T, unsigned base, double2* tile) {
    const DmRegRound& R = P.r[r];
    const bool first = r == 0, last = r == P.n_rounds - 1;
    int e = 0;
    unsigned g = base;
#pragma unroll
    for (int k = 0; k < LB; ++k) {
        const int m = -((tid >> k) & 1);
        e |= m & R.emask[k];
        g |= (unsigned)m & R.gmask[k];
    }
    const int pe = dm_phys(e);
    const int hp0 = R.hp[0], hp1 = R.hp[1], hp2 = R.hp[2], hp3 = R.hp[3];
    const unsigned hg0 = R.hg[0], hg1 = R.hg[1], hg2 = R.hg[2], hg3 = R.hg[3];
#define DMR_G(j) (g | (((j) & 1) ? hg0 : 0u) | (((j) & 2) ? hg1 : 0u) | (((j) & 4) ? hg2 : 0u) | (((j) & 8) ? hg3 : 0u))
#define DMR_S(j) (pe ^ (((j) & 1) ? hp0 : 0) ^ (((j) & 2) ? hp1 : 0) ^ (((j) & 4) ? hp2 : 0) ^ (((j) & 8) ? hp3 : 0))
    double2 v[16];
    if (first) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = DMR_LOAD_STREAM(rho + DMR_G(j));
        if (P.has_diag) {
            const unsigned rmask = (1u << P.n) - 1u;
            const bool rcrc = hg0 <= rmask && hg1 > rmask && hg2 <= rmask && hg3 > rmask;
            if (rcrc) {
                // the standard round: two row bits and two column bits in registers -> four table entries each (rows first,
                // then columns: at most four entries live beside the 16 elements)
                {
                    const unsigned gr = g & rmask;
                    double2 tr[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) tr[k] = DMR_LOAD_TABLE(T + (gr | ((k & 1) ? hg0 : 0u) | ((k & 2) ? hg2 : 0u)));
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = cmul(v[j], tr[(j & 1) | ((j >> 1) & 2)]);
                }
                {
                    const unsigned gc = g >> P.n, hc1 = hg1 >> P.n, hc3 = hg3 >> P.n;
                    double2 tc[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double2 c = DMR_LOAD_TABLE(T + (gc | ((k & 1) ? hc1 : 0u) | ((k & 2) ? hc3 : 0u)));
                        tc[k] = make_double2(c.x, -c.y);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = cmul(v[j], tc[((j >> 1) & 1) | ((j >> 2) & 2)]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const unsigned gj = DMR_G(j);
                    const double2 a = DMR_LOAD_TABLE(T + (gj & rmask)), c = DMR_LOAD_TABLE(T + (gj >> P.n));
                    v[j] = cmul(cmul(v[j], a), make_double2(c.x, -c.y));
                }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tile[DMR_S(j)];
    }
    dmr_apply_qubit<0, 1>(v, R.op[0]);
    if (R.nop > 1) dmr_apply_qubit<2, 3>(v, R.op[1]);
    if (last) {
#pragma unroll
        for (int j = 0; j < 16; ++j) DMR_STORE_STREAM(rho + DMR_G(j), v[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) tile[DMR_S(j)] = v[j];
    }
#undef DMR_G
#undef DMR_S
}

// ------------------------------------------------------------------------------------ host: planner
struct DmStep {
    int kind;                        // 0: build the phase table from D; 1: apply the table elementwise; 2: k_dm_tile; 3: k_dm_reg
    DmDiagTerms D;
    DmTilePass P;
    DmRegPass R;
};

struct DmPlanOptions {
    bool reg_passes = true;          // register-resident passes for tiles of 2^12 / 2^13 elements (else k_dm_tile everywhere)
    bool wide13 = true;              // groups without qubit 0: row bit 0 as a passive bit in a 2^13 tile (32 B runs)
};

static inline DmRegOp dm_reg_op(double theta, double px, double py, double pz) {
    const double c = cos(0.5 * theta), s = sin(0.5 * theta);
    double dA = 1.0 - px - py, dB = px + py, oA = 1.0 - px - py - 2.0 * pz, oB = px - py;
    DmRegOp o;
    double scale;
    if (fabs(c) >= fabs(s)) {
        o.t = s / c;
        scale = c * c;
    } else {
        // RX(theta) = -i s X (1 - i t' X), t' = -c / s: rotate by t', then X rho X = swap of the block's partners
        o.t = -c / s;
        scale = s * s;
        const double u = dA; dA = dB; dB = u;
        const double w = oA; oA = oB; oB = w;
    }
    o.A = scale * dA; o.B = scale * dB; o.C = scale * oA; o.D = scale * oB;
    return o;
}

// image of tile-local bit b in the low three bits of its shared-memory position
static inline int dm_phys_low3(int b) { return dm_phys(1 << b) & 7; }

// thread-index bits of a round: lanes 0-2 get three free bits whose swizzle images are independent (conflict-free 16 B
// accesses), preferring low bits (low bits in a warp = long runs in HBM); the rest ascending
static inline void dm_choose_thread_bits(DmRegRound& R, int tile_bits) {
    int freeb[13], nf = 0;
    for (int b = 0; b < tile_bits; ++b)
        if (b != R.hb[0] && b != R.hb[1] && b != R.hb[2] && b != R.hb[3]) freeb[nf++] = b;
    int pick[3] = {-1, -1, -1};
    bool found = false;
    for (int i = 0; i < nf && !found; ++i)
        for (int j = i + 1; j < nf && !found; ++j)
            for (int k = j + 1; k < nf && !found; ++k) {
                const int a = dm_phys_low3(freeb[i]), b = dm_phys_low3(freeb[j]), c = dm_phys_low3(freeb[k]);
                // independent over GF(2): no non-empty subset XORs to zero
                if (a && b && c && (a ^ b) && (a ^ c) && (b ^ c) && (a ^ b ^ c)) {
                    pick[0] = freeb[i]; pick[1] = freeb[j]; pick[2] = freeb[k];
                    found = true;
                }
            }
    int n = 0;
    if (found)
        for (int k = 0; k < 3; ++k) R.tbit[n++] = pick[k];
    for (int i = 0; i < nf; ++i)
        if (!found || (freeb[i] != pick[0] && freeb[i] != pick[1] && freeb[i] != pick[2])) R.tbit[n++] = freeb[i];
}

// derived fields of a round (after hb / tbit are chosen)
static inline void dm_finish_round(DmRegRound& R, const int* tb, int tile_bits) {
    for (int k = 0; k < 9; ++k) { R.emask[k] = 0; R.gmask[k] = 0; }
    for (int k = 0; k < tile_bits - 4; ++k) {
        R.emask[k] = 1 << R.tbit[k];
        R.gmask[k] = 1u << tb[R.tbit[k]];
    }
    for (int k = 0; k < 4; ++k) {
        R.hp[k] = dm_phys(1 << R.hb[k]);
        R.hg[k] = 1u << tb[R.hb[k]];
    }
}

struct DmQOpHost { bool used; double theta, px, py, pz; };

// tile bits of a group: optional passive row bit 0, the row and column bits of the group's qubits, spare positions filled with
// the lowest unused bits; CTA index -> base segments
static inline bool dm_tile_geometry(int n, int tile_bits, bool passive, const int* grp, int ng, int* tb, int& seg_n, int* seg_src,
                                    int* seg_len, int* seg_dst) {
    u64 used = 0;
    for (int k = 0; k < ng; ++k) used |= (1ull << grp[k]) | (1ull << (grp[k] + n));
    if (passive) used |= 1ull;
    for (int b = 0; b < 2 * n && __builtin_popcountll(used) < tile_bits; ++b) used |= 1ull << b;
    int l = 0;
    for (int b = 0; b < 2 * n; ++b)
        if ((used >> b) & 1ull) tb[l++] = b;
    seg_n = 0;
    int src = 0, p2 = 0;
    while (p2 < 2 * n) {
        if ((used >> p2) & 1ull) { ++p2; continue; }
        int len = 0;
        while (p2 + len < 2 * n && !((used >> (p2 + len)) & 1ull)) ++len;
        if (seg_n >= 8) return false;
        seg_src[seg_n] = src; seg_len[seg_n] = len; seg_dst[seg_n] = p2; ++seg_n;
        src += len; p2 += len;
    }
    return true;
}

static inline bool dm_plan(int n, int n_seg, const int32_t* seg_type, const int32_t* seg_off, const int32_t* q0,
                           const int32_t* q1, const double* val, const double* probs, const DmPlanOptions& opt,
                           std::vector<DmStep>& steps, std::string& err) {
    bool diag_pending = false;
    const int TB = (2 * n < 12) ? 2 * n : 12;
    int i = 0;
    while (i < n_seg) {
        if (seg_type[i] == 1) {
            if (diag_pending) {                          // two diagonal segments in a row: flush the first
                DmStep S;
                memset(&S, 0, sizeof(S));
                S.kind = 1;
                steps.push_back(S);
            }
            DmStep S;
            memset(&S, 0, sizeof(S));
            S.kind = 0;
            DmDiagTerms& D = S.D;
            for (int k = seg_off[i]; k < seg_off[i + 1]; ++k) {
                if (q0[k] < 0 || q0[k] >= n || q1[k] >= n) { err = "diagonal term: qubit out of range"; return false; }
                if (q1[k] < 0) {
                    if (D.n1 >= 16) { err = "too many one-body terms"; return false; }
                    D.q1[D.n1] = q0[k]; D.a[D.n1] = val[k]; ++D.n1;
                } else {
                    if (D.n2 >= DTC_MAXT) { err = "too many two-body terms in one segment"; return false; }
                    D.qi[D.n2] = q0[k]; D.qj[D.n2] = q1[k]; D.b[D.n2] = val[k]; ++D.n2;
                }
            }
            steps.push_back(S);
            diag_pending = true;
            ++i;
            continue;
        }
        // a layer of qubit operations: [rotations] [channels]
        DmQOpHost ops[16];
        for (int q = 0; q < 16; ++q) ops[q] = DmQOpHost{false, 0.0, 0.0, 0.0, 0.0};
        int j = i;
        if (seg_type[j] == 0) {
            for (int k = seg_off[j]; k < seg_off[j + 1]; ++k) {
                if (q0[k] < 0 || q0[k] >= n || ops[q0[k]].used) { err = "rotation segment: bad or repeated qubit"; return false; }
                ops[q0[k]].used = true;
                ops[q0[k]].theta = val[k];
            }
            ++j;
        }
        if (j < n_seg && seg_type[j] == 2) {
            bool seen[16] = {false};
            for (int k = seg_off[j]; k < seg_off[j + 1]; ++k) {
                if (q0[k] < 0 || q0[k] >= n || seen[q0[k]]) { err = "channel segment: bad or repeated qubit"; return false; }
                seen[q0[k]] = true;
                DmQOpHost& o = ops[q0[k]];
                o.used = true;
                o.px = probs[3 * k]; o.py = probs[3 * k + 1]; o.pz = probs[3 * k + 2];
            }
            ++j;
        } else if (j == i) {
            err = "unknown segment type";
            return false;
        }
        i = j;
        int todo[16], nt = 0;
        for (int q = 0; q < n; ++q)
            if (ops[q].used) todo[nt++] = q;
        int pos = 0;
        while (pos < nt) {
            // a group that holds qubit 0 (or a register that fits one tile): up to six qubits in a 2^12 tile, row bit 0 among
            // them; any other group: row bit 0 as a passive bit (32 B runs = whole sectors) + up to six qubits in a 2^13 tile
            const bool low_group = todo[pos] == 0 || 2 * n <= 12 || !opt.wide13;
            const int passive = low_group ? 0 : 1;
            const int TBg = low_group ? TB : ((2 * n < 13) ? 2 * n : 13);
            int cap = (TBg - passive) / 2;
            if (cap > DTC_DM_MAXQ) cap = DTC_DM_MAXQ;
            int grp[DTC_DM_MAXQ], ng = 0;
            while (pos < nt && ng < cap) grp[ng++] = todo[pos++];
            DmStep S;
            memset(&S, 0, sizeof(S));
            const bool reg = opt.reg_passes && TBg >= 12;
            S.kind = reg ? 3 : 2;
            if (!reg) {
                DmTilePass& P = S.P;
                P.n = n; P.tile_bits = TBg; P.nq = ng;
                if (!dm_tile_geometry(n, TBg, passive != 0, grp, ng, P.tb, P.seg_n, P.seg_src, P.seg_len, P.seg_dst)) {
                    err = "internal: too many index segments";
                    return false;
                }
                for (int k = 0; k < ng; ++k) {
                    DmQubitOp& Q = P.q[k];
                    for (int m = 0; m < TBg; ++m) {
                        if (P.tb[m] == grp[k]) Q.lr = m;
                        if (P.tb[m] == grp[k] + n) Q.lc = m;
                    }
                    const DmQOpHost& o = ops[grp[k]];
                    Q.c = cos(0.5 * o.theta); Q.s = sin(0.5 * o.theta);
                    Q.dA = 1.0 - o.px - o.py; Q.dB = o.px + o.py;
                    Q.oA = 1.0 - o.px - o.py - 2.0 * o.pz; Q.oB = o.px - o.py;
                }
                P.has_diag = diag_pending ? 1 : 0;
            } else {
                DmRegPass& P = S.R;
                P.n = n; P.tile_bits = TBg;
                if (!dm_tile_geometry(n, TBg, passive != 0, grp, ng, P.tb, P.seg_n, P.seg_src, P.seg_len, P.seg_dst)) {
                    err = "internal: too many index segments";
                    return false;
                }
                // pairs from the top; the round that loads from HBM takes the top pair (low bits stay with the lanes), the
                // round that stores to HBM the middle one, the lowest pair (or single qubit) is the shared-memory-only round
                int pr[3][2], np = 0;
                for (int k = ng; k > 0; k -= 2) {
                    pr[np][0] = k >= 2 ? grp[k - 2] : grp[k - 1];
                    pr[np][1] = k >= 2 ? grp[k - 1] : -1;
                    ++np;
                }
                int order[3] = {0, 1, 2};
                if (np == 3) { order[0] = 0; order[1] = 2; order[2] = 1; }
                P.n_rounds = np;
                for (int r = 0; r < np; ++r) {
                    DmRegRound& R = P.r[r];
                    const int* pq = pr[order[r]];
                    auto local_bit = [&](int gb) {
                        for (int m = 0; m < TBg; ++m)
                            if (P.tb[m] == gb) return m;
                        return -1;
                    };
                    R.hb[0] = local_bit(pq[0]); R.hb[1] = local_bit(pq[0] + n);
                    R.op[0] = dm_reg_op(ops[pq[0]].theta, ops[pq[0]].px, ops[pq[0]].py, ops[pq[0]].pz);
                    if (pq[1] >= 0) {
                        R.nop = 2;
                        R.hb[2] = local_bit(pq[1]); R.hb[3] = local_bit(pq[1] + n);
                        R.op[1] = dm_reg_op(ops[pq[1]].theta, ops[pq[1]].px, ops[pq[1]].py, ops[pq[1]].pz);
                    } else {
                        R.nop = 1;
                        int k = 2;
                        for (int b = TBg - 1; b >= 0 && k < 4; --b)
                            if (b != R.hb[0] && b != R.hb[1]) R.hb[k++] = b;
                    }
                    dm_choose_thread_bits(R, TBg);
                    dm_finish_round(R, P.tb, TBg);
                }
                P.has_diag = diag_pending ? 1 : 0;
            }
            diag_pending = false;
            steps.push_back(S);
        }
    }
    if (diag_pending) {
        DmStep S;
        memset(&S, 0, sizeof(S));
        S.kind = 1;
        steps.push_back(S);
    }
    return true;
}
