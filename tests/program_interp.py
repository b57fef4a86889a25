"""numpy interpreter of a compiled dtcsim Program (TEST INFRASTRUCTURE).

Executable specification of the semantics the CUDA library implements for the event stream of
plan.py: Pauli-frame walk -> per-(layer, trajectory) sign masks -> tan-form RX rotations with the
cos scale folded into the layer constant -> diagonal layers -> frame materialisation.  Used by the
CPU tests to check the planner against the gate-by-gate oracle without a GPU.
"""
import math

import numpy as np

from oracle import philox

EV_ROT, EV_D1, EV_D2, EV_NOISE, EV_D2C = 0, 1, 2, 3, 4


def build_layers(prog):
    """Host-side table prep (mirrors dtc_program_finalize)."""
    n, M = prog.n, prog.n_layers
    L = [dict(tan=np.zeros(n), scale=1.0, a=np.zeros((2, n)), terms=[]) for _ in range(M)]
    nm = prog.n_main
    ev = prog.arrays()
    k_of = np.zeros(len(ev["type"]), dtype=np.int64)
    for e in range(len(ev["type"])):
        t, lay, q0, q1, slot, val = (int(ev["type"][e]), int(ev["layer"][e]), int(ev["q0"][e]),
                                     int(ev["q1"][e]), int(ev["slot"][e]), float(ev["val"][e]))
        if t == EV_ROT:
            k = int(np.rint(val / math.pi))
            th = val - k * math.pi
            if math.tan(th / 2) != 0.0:
                assert L[lay]["tan"][q0] == 0.0, "two rotations on one qubit in one layer"
                L[lay]["tan"][q0] = math.tan(th / 2)
                L[lay]["scale"] *= math.cos(th / 2)
                k_of[e] = k % 4
            else:
                k_of[e] = (k % 4) + 4                    # +4 marks a pure Pauli RX(k pi): no sign bit
        elif t == EV_D1:
            L[lay]["a"][slot, q0] = val
        elif t in (EV_D2, EV_D2C):
            assert slot == len(L[lay]["terms"])
            L[lay]["terms"].append((q0, q1 if t == EV_D2 else -1, val))     # -1: partner is |0> in psi'
    return L, k_of


def frame_walk(prog, k_of, seed, trajs):
    """Returns masks[layer][4][traj] (uint64) and final frame (fx, fz, ph)."""
    ev = prog.arrays()
    trajs = np.asarray(trajs, dtype=np.uint64)
    T = len(trajs)
    masks = np.zeros((prog.n_layers, 4, T), dtype=np.uint64)
    fx = np.zeros(T, dtype=np.uint64)
    fz = np.zeros(T, dtype=np.uint64)
    ph = np.zeros(T, dtype=np.int64)
    one = np.uint64(1)
    for e in range(len(ev["type"])):
        t, lay, q0, q1, slot = (int(ev["type"][e]), int(ev["layer"][e]), int(ev["q0"][e]),
                                int(ev["q1"][e]), int(ev["slot"][e]))
        b0 = np.uint64(q0)
        if t == EV_ROT:
            if k_of[e] < 4:
                masks[lay, 0] |= ((fz >> b0) & one) << b0
            k = int(k_of[e]) % 4
            if k & 1:
                fx ^= one << b0
            ph += 3 * k
        elif t == EV_D1:
            masks[lay, 1 + slot] |= ((fx >> b0) & one) << b0
        elif t in (EV_D2, EV_D2C):
            b1 = np.uint64(q1)
            masks[lay, 3] |= (((fx >> b0) ^ (fx >> b1)) & one) << np.uint64(slot)
        elif t == EV_NOISE:
            px, py, pz = ev["probs"][e]
            u = philox.uniform(seed, slot, philox.STREAM_NOISE, trajs)
            code = philox.pauli_from_uniform(u, px, py, pz)
            fxq = ((fx >> b0) & one).astype(np.int64)
            isx, isy, isz = code == 1, code == 2, code == 3
            ph += np.where(isy, 1 + 2 * fxq, 0) + np.where(isz, 2 * fxq, 0)
            fx ^= np.where(isx | isy, one << b0, np.uint64(0))
            fz ^= np.where(isy | isz, one << b0, np.uint64(0))
    return masks, fx, fz, ph % 4


def _bits(n, q):
    return ((np.arange(1 << n, dtype=np.int64) >> q) & 1)


def apply_rot_tan(psi, q, t):
    """psi [T, 2^n]; tan-form RX: out0 = x0 - i t x1, out1 = x1 - i t x0 with per-trajectory t [T]."""
    T = psi.shape[0]
    v = psi.reshape(T, -1, 2, 1 << q)
    tt = (1j * t)[:, None, None]
    x0 = v[:, :, 0, :].copy()
    x1 = v[:, :, 1, :].copy()
    v[:, :, 0, :] = x0 - tt * x1
    v[:, :, 1, :] = x1 - tt * x0
    return psi


def diag_phase(layer, const, n, m1a, m1b, m2):
    """phase[T, 2^n] of one D layer for mask vectors (uint64 [T])."""
    T = len(m1a)
    ph = np.full((T, 1 << n), const, dtype=np.complex128)
    one = np.uint64(1)
    for q in range(n):
        z = 1.0 - 2.0 * _bits(n, q)
        for slot, m in ((0, m1a), (1, m1b)):
            a = layer["a"][slot, q]
            if a != 0.0:
                sg = 1.0 - 2.0 * ((m >> np.uint64(q)) & one).astype(np.float64)
                ph *= math.cos(a / 2) - 1j * math.sin(a / 2) * sg[:, None] * z[None, :]
    for k, (i, j, b) in enumerate(layer["terms"]):
        zz = (1.0 - 2.0 * _bits(n, i)) * ((1.0 - 2.0 * _bits(n, j)) if j >= 0 else 1.0)
        sg = 1.0 - 2.0 * ((m2 >> np.uint64(k)) & one).astype(np.float64)
        ph *= math.cos(b / 2) - 1j * math.sin(b / 2) * sg[:, None] * zz[None, :]
    return ph


def run(prog, seed=0, trajs=(0,), init_index=0, materialize=True):
    """Execute the program for the given trajectory ids; returns psi_true [T, 2^n] (or psi', frame)."""
    n = prog.n_main                                   # < prog.n when the read-out is factorised
    layers, k_of = build_layers(prog)
    layers = layers[:prog.n_exec_layers]
    masks, fx, fz, ph = frame_walk(prog, k_of, seed, trajs)
    T = len(trajs)
    psi = np.zeros((T, 1 << n), dtype=np.complex128)
    psi[:, init_index] = 1.0
    one = np.uint64(1)
    for j, lay in enumerate(layers):
        for q in range(n):
            if lay["tan"][q] != 0.0:
                sg = 1.0 - 2.0 * ((masks[j, 0] >> np.uint64(q)) & one).astype(np.float64)
                psi = apply_rot_tan(psi, q, lay["tan"][q] * sg)
        const = lay["scale"] * (np.exp(1j * prog.global_phase) if j == 0 else 1.0)
        psi = psi * diag_phase(lay, const, n, masks[j, 1], masks[j, 2], masks[j, 3])
    if not materialize:
        return psi, (fx, fz, ph)
    return materialize_frame(psi, n, fx, fz, ph)


def materialize_frame(psi, n, fx, fz, ph):
    """psi_true(y) = i^ph (-1)^{popc((y^fx)&fz)} psi'(y ^ fx)."""
    out = np.empty_like(psi)
    y = np.arange(1 << n, dtype=np.uint64)
    for r in range(psi.shape[0]):
        src = y ^ fx[r]
        par = np.zeros(1 << n, dtype=np.int64)
        m = src & fz[r]
        for b in range(n):
            par ^= ((m >> np.uint64(b)) & np.uint64(1)).astype(np.int64)
        out[r] = (1j ** int(ph[r])) * (1 - 2 * par) * psi[r, src.astype(np.int64)]
    return out


def run_dm(prog, n):
    """Execute prog.dm_segments on a density matrix (vectorised rho as a 2n-qubit state)."""
    d = 1 << n
    rho = np.zeros((1, d * d), dtype=np.complex128)     # index = row + d*col (row bits low)
    rho[0, 0] = 1.0
    for seg in prog.dm_segments:
        if seg[0] == "R":
            for q, th in seg[1]:
                k = int(np.rint(th / math.pi))
                thp = th - k * math.pi
                # (-iX)^k on rows, conj on columns: phases cancel, X^k on both
                if k & 1:
                    idx = np.arange(d * d, dtype=np.int64) ^ (1 << q) ^ (1 << (q + n))
                    rho = rho[:, idx]
                t = math.tan(thp / 2)
                rho = apply_rot_tan(rho, q, np.array([t]))
                rho = apply_rot_tan(rho, q + n, np.array([-t]))       # conj(RX(t)) = RX(-t)
                rho *= math.cos(thp / 2) ** 2
        elif seg[0] == "D":
            _, d1, d2 = seg
            ph = np.ones(d * d, dtype=np.complex128)
            for q, a in d1.items():
                zr = 1.0 - 2.0 * _bits(2 * n, q)
                zc = 1.0 - 2.0 * _bits(2 * n, q + n)
                ph *= np.exp(-0.5j * a * (zr - zc))
            for (i, j), b in d2.items():
                zr = (1.0 - 2.0 * _bits(2 * n, i)) * (1.0 - 2.0 * _bits(2 * n, j))
                zc = (1.0 - 2.0 * _bits(2 * n, i + n)) * (1.0 - 2.0 * _bits(2 * n, j + n))
                ph *= np.exp(-0.5j * b * (zr - zc))
            rho = rho * ph[None, :]
        elif seg[0] == "K":
            # general single-qubit channel: 4 x 4 superoperator on the (row bit q, column bit q) block, index = row + 2 col
            for q, S in seg[1]:
                idx = np.arange(d * d, dtype=np.int64)
                base = idx & ~((1 << q) | (1 << (q + n)))
                blk = ((idx >> q) & 1) | (((idx >> (q + n)) & 1) << 1)
                src = [base | ((j & 1) << q) | ((j >> 1) << (q + n)) for j in range(4)]
                S = np.asarray(S, dtype=np.complex128)
                new = np.zeros(d * d, dtype=np.complex128)
                for j in range(4):
                    new += S[blk, j] * rho[0, src[j]]
                rho = new[None, :]
        elif seg[0] == "N":
            for q, (px, py, pz) in seg[1]:
                idx = np.arange(d * d, dtype=np.int64)
                partner = idx ^ (1 << q) ^ (1 << (q + n))
                s = 1.0 - 2.0 * (_bits(2 * n, q) ^ _bits(2 * n, q + n))
                e = rho[0]
                f = rho[0, partner]
                rho = ((1 - px - py - pz) * e + px * f + py * s * f + pz * s * e)[None, :]
    return rho[0].reshape(d, d).T            # [row, col]


def to_circuit_order(psi, prog):
    """psi [..., 2^n] in the program's internal bit order -> compacted circuit qubit order."""
    n = prog.n
    if list(prog.order) == list(range(n)):
        return psi
    lead = psi.shape[:-1]
    v = psi.reshape(lead + (2,) * n)                    # axis (len(lead) + n-1-b) <-> bit b
    nl = len(lead)
    # new bit c (compacted qubit c) takes old bit prog.bit_of[c]
    axes = list(range(nl)) + [nl + n - 1 - prog.bit_of[n - 1 - k] for k in range(n)]
    return np.ascontiguousarray(np.transpose(v, axes)).reshape(lead + (1 << n,))


def dm_to_circuit_order(rho, prog):
    n = prog.n
    r = to_circuit_order(rho, prog)                     # columns
    return to_circuit_order(r.T, prog).T                # rows
