/* dtc_oracle.c -- plain-C restatement of the gate-by-gate statevector path (TEST INFRASTRUCTURE).
 *
 * Same algorithm as oracle/oracle.py (which documents the reference call sites it follows:
 * backend.run(circ, shots) fast.py:211, depolarizing noise fast.py:85-86), written in C + OpenMP so
 * that (a) full-size states (n = 21) can be checked amplitude by amplitude and (b) bench.py has a
 * CPU baseline that executes the circuit the way Aer does: one pass over the state per transpiled
 * gate, one trajectory per shot, complex128.  Never linked into the product library.
 *
 * Gate definitions: Qiskit circuit library (u3/u2/u1/rz as 2x2 matrices supplied by the caller,
 * cx as an index permutation); little-endian qubit order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

static inline cplx cmul(cplx a, cplx b) { cplx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cplx cadd(cplx a, cplx b) { cplx r = {a.re + b.re, a.im + b.im}; return r; }

/* 2x2 matrix u (row-major, 4 complex) on qubit q */
void orc_apply_1q(cplx* psi, int n, int q, const cplx* u) {
    const int64_t half = (int64_t)1 << (n - 1);
    const int64_t low = ((int64_t)1 << q) - 1;
    const int64_t bit = (int64_t)1 << q;
    const int diag = (u[1].re == 0 && u[1].im == 0 && u[2].re == 0 && u[2].im == 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < half; ++i) {
        const int64_t i0 = ((i & ~low) << 1) | (i & low);
        const int64_t i1 = i0 | bit;
        const cplx a = psi[i0], b = psi[i1];
        if (diag) {
            psi[i0] = cmul(u[0], a);
            psi[i1] = cmul(u[3], b);
        } else {
            psi[i0] = cadd(cmul(u[0], a), cmul(u[1], b));
            psi[i1] = cadd(cmul(u[2], a), cmul(u[3], b));
        }
    }
}

void orc_apply_cx(cplx* psi, int n, int c, int t) {
    const int64_t N = (int64_t)1 << n;
    const int64_t cb = (int64_t)1 << c, tb = (int64_t)1 << t;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        if ((i & cb) && !(i & tb)) {
            const cplx a = psi[i];
            psi[i] = psi[i | tb];
            psi[i | tb] = a;
        }
    }
}

/* Run a whole circuit on one state.  kind: 0 = 1q matrix (mats[8*i..]), 1 = cx(q0 -> q1), 2 = no-op.
 * pauli[i] in {0,1,2,3} = I,X,Y,Z applied on q0 after op i (the sampled noise). */
void orc_run(cplx* psi, int n, int64_t n_ops, const int32_t* kind, const int32_t* q0, const int32_t* q1,
             const double* mats, const uint8_t* pauli) {
    static const cplx PX[4] = {{0, 0}, {1, 0}, {1, 0}, {0, 0}};
    static const cplx PY[4] = {{0, 0}, {0, -1}, {0, 1}, {0, 0}};
    static const cplx PZ[4] = {{1, 0}, {0, 0}, {0, 0}, {-1, 0}};
    for (int64_t i = 0; i < n_ops; ++i) {
        if (kind[i] == 0) orc_apply_1q(psi, n, q0[i], (const cplx*)(mats + 8 * i));
        else if (kind[i] == 1) orc_apply_cx(psi, n, q0[i], q1[i]);
        if (pauli && pauli[i]) {
            const cplx* P = pauli[i] == 1 ? PX : (pauli[i] == 2 ? PY : PZ);
            orc_apply_1q(psi, n, q0[i], P);
        }
    }
}

void orc_init(cplx* psi, int n, int64_t index) {
    const int64_t N = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) { psi[i].re = 0; psi[i].im = 0; }
    psi[index].re = 1.0;
}

/* P(bit q = 1) */
double orc_prob1(const cplx* psi, int n, int q) {
    const int64_t N = (int64_t)1 << n;
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < N; ++i)
        if ((i >> q) & 1) s += psi[i].re * psi[i].re + psi[i].im * psi[i].im;
    return s;
}

#ifdef _OPENMP
#include <omp.h>
int orc_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
#else
int orc_threads(void) { return 1; }
void orc_set_threads(int n) { (void)n; }
#endif
