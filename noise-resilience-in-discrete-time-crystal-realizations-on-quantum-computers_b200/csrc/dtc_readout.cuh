// dtc_readout.cuh -- read-out of a factorised circuit (plan.compile_circuit(optimize=True)), per trajectory.
//
// For the reference's Hadamard-test circuits (fast.py:125-147) the ancilla never enters the big register
// (signal = (1-p)^6 s_q <Z_q>): the device evolves the L-qubit state, dtc_rdm reduces it to
// the density matrix of the partner qubit(s), and the few events after the last Floquet period act on
// <= 3 qubits.  This applies those "small" events -- with the trajectory's own sign masks from the frame
// walk -- to rho = rdm (x) |0><0| and returns the outcome probabilities of the measured bits (what
// AerSimulator's measure sampling consumes, fast.py:211).  One thread per trajectory; compiles under nvcc
// and g++ (tests/emul runs it against the numpy restatement tests/readout_ref.py).
#pragma once
#include "dtc_hd.cuh"

#define DTC_SMALL_MAXQ 3
#define DTC_SMALL_DIM (1 << DTC_SMALL_MAXQ)

struct DtcSmallPlan {
    int nq, n_reg, m, pad_;
    int bits[DTC_SMALL_MAXQ];        // internal bit index of small position p (register bits first, eliminated bits after)
    int meas_pos[DTC_SMALL_MAXQ];    // small position of measured bit i
    int meas_bit[DTC_SMALL_MAXQ];    // internal bit index of measured bit i (frame x-mask flips its outcome)
};

DTC_HD int small_pos(const DtcSmallPlan& S, int bit) {
    for (int p = 0; p < S.nq; ++p)
        if (S.bits[p] == bit) return p;
    return -1;
}

// masks: pointer to this trajectory's entry of mask row 0 (rows are mstride apart); rdm: [2^n_reg][2^n_reg]
DTC_HD void small_readout_traj(const DtcSmallPlan& S, const DtcEvent* ev, const long long* idx, long long n_small,
                               const double2* rdm, const u64* masks, long long mstride, u64 fx, double* probs) {
    const int d = 1 << S.nq, dr = 1 << S.n_reg;
    double2 rho[DTC_SMALL_DIM][DTC_SMALL_DIM];
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) rho[i][j] = (i < dr && j < dr) ? rdm[i * dr + j] : make_double2(0.0, 0.0);
    for (long long e = 0; e < n_small; ++e) {
        const DtcEvent E = ev[idx[e]];
        if (E.type == DTC_EVT_NOISE) continue;                   // already folded into the frame
        const int p0 = small_pos(S, E.q0);
        if (p0 < 0) continue;
        if (E.type == DTC_EVT_ROT) {
            if (E.slot == 1) continue;                           // pure Pauli rotation: frame update only
            const double theta = E.c0;
            const double thp = theta - nearbyint(theta / M_PI) * M_PI;
            const bool neg = (masks[(long long)(E.layer * 4 + 0) * mstride] >> E.q0) & 1ull;
            const double c = cos(0.5 * thp), s = neg ? -sin(0.5 * thp) : sin(0.5 * thp);
            const int bit = 1 << p0;
            for (int i = 0; i < d; ++i) {                        // rows: rho <- U rho, U = [[c, -is], [-is, c]]
                if (i & bit) continue;
                for (int j = 0; j < d; ++j) {
                    const double2 a = rho[i][j], b = rho[i | bit][j];
                    rho[i][j] = make_double2(c * a.x + s * b.y, c * a.y - s * b.x);
                    rho[i | bit][j] = make_double2(c * b.x + s * a.y, c * b.y - s * a.x);
                }
            }
            for (int j = 0; j < d; ++j) {                        // columns: rho <- rho U^dagger
                if (j & bit) continue;
                for (int i = 0; i < d; ++i) {
                    const double2 a = rho[i][j], b = rho[i][j | bit];
                    rho[i][j] = make_double2(c * a.x - s * b.y, c * a.y + s * b.x);
                    rho[i][j | bit] = make_double2(c * b.x - s * a.y, c * b.y + s * a.x);
                }
            }
            continue;
        }
        bool neg;
        int zmask;                                               // z_i = (-1)^{popc(i & zmask)}
        if (E.type == DTC_EVT_D1) {
            neg = (masks[(long long)(E.layer * 4 + 1 + E.slot) * mstride] >> E.q0) & 1ull;
            zmask = 1 << p0;
        } else {
            neg = (masks[(long long)(E.layer * 4 + 3) * mstride] >> E.slot) & 1ull;
            zmask = 1 << p0;
            if (E.type == DTC_EVT_D2) {
                const int p1 = small_pos(S, E.q1);
                if (p1 < 0) continue;
                zmask |= 1 << p1;
            }
        }
        const double half = neg ? -0.5 * E.c0 : 0.5 * E.c0;
        // ph_i = exp(-i half z_i);  rho_ij *= ph_i conj(ph_j) = exp(-i half (z_i - z_j)): identity when z_i = z_j
        const double2 up = make_double2(cos(2.0 * half), -sin(2.0 * half));      // z_i = +1, z_j = -1
        for (int i = 0; i < d; ++i) {
            const int zi = DTC_POPC64((u64)(i & zmask)) & 1;
            for (int j = 0; j < d; ++j) {
                const int zj = DTC_POPC64((u64)(j & zmask)) & 1;
                if (zi == zj) continue;
                rho[i][j] = cmul(rho[i][j], zi ? make_double2(up.x, -up.y) : up);
            }
        }
    }
    const int nb = 1 << S.m;
    for (int b = 0; b < nb; ++b) probs[b] = 0.0;
    int flip = 0;
    for (int i = 0; i < S.m; ++i) flip |= (int)((fx >> S.meas_bit[i]) & 1ull) << i;
    for (int v = 0; v < d; ++v) {
        int col = 0;
        for (int i = 0; i < S.m; ++i) col |= ((v >> S.meas_pos[i]) & 1) << i;
        probs[col ^ flip] += rho[v][v].x;
    }
}
