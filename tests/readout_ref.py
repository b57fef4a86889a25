"""Host half of the read-out factorisation (plan.compile_circuit(optimize=True)).

For the reference's Hadamard-test circuits (fast.py:125-147) the ancilla never has to live in the big
register: before the final ``cz(q+1, 0)`` it only meets system qubits that are still classical, so
its state is a single-qubit vector, and the few operations after the last Floquet period act on the
ancilla and site q alone (signal = (1-p)^6 s_q <Z_q>).  The device evolves the
L-qubit state, reduces it to the density matrix of the partner qubit(s) with ``dtc_rdm`` and this
module applies the remaining "small" events -- with the *same* sampled Pauli frames (sign masks
from the device's frame walk) -- to a <= 3-qubit density matrix per trajectory, vectorised with
numpy.  The work is O(trajectories * events * 64) flops: read-out glue, not the hot path.
"""
import math

import numpy as np

from dtcsim.plan import EV_D1, EV_D2, EV_D2C, EV_NOISE, EV_ROT


def _full_1q(U, p, nq):
    """[T,2,2] single-qubit matrices on position p of an nq-qubit space -> [T, 2^nq, 2^nq]."""
    T = U.shape[0]
    lo, hi = 1 << p, 1 << (nq - 1 - p)
    full = np.einsum("ab,tij,cd->taicbjd", np.eye(hi), U, np.eye(lo))
    return full.reshape(T, hi * 2 * lo, hi * 2 * lo)


def simulate_small(prog, rdm, masks, fx, first_mask_layer=0):
    """Apply prog.small's events to rho = rdm (x) |0><0|_E for every trajectory.

    rdm:   [T, 2^k, 2^k] complex reduced density matrices of psi' on prog.small['reg_bits'] (bit i of
           the index = reg_bits[i]);  masks: [n_layers - first_mask_layer, 4, T] uint64 sign masks;
    fx:    [T] uint64 final frame x-masks.  Returns probabilities [T, 2^m] over the measured qubits
    (column bit i = i-th entry of prog.measures), frame flips applied.
    """
    sm = prog.small
    bits = list(sm["reg_bits"]) + list(sm["elim_bits"])
    pos = {b: i for i, b in enumerate(bits)}
    nq = len(bits)
    d = 1 << nq
    T = rdm.shape[0]
    dr = 1 << len(sm["reg_bits"])
    rho = np.zeros((T, d, d), dtype=np.complex128)
    rho[:, :dr, :dr] = rdm
    idx = np.arange(d)
    zbit = [1.0 - 2.0 * ((idx >> p) & 1) for p in range(nq)]
    one = np.uint64(1)
    ev = prog.arrays()
    for e in sm["events"]:
        typ, layer, q0, q1, slot, val = (int(ev["type"][e]), int(ev["layer"][e]) - first_mask_layer, int(ev["q0"][e]),
                                         int(ev["q1"][e]), int(ev["slot"][e]), float(ev["val"][e]))
        if typ == EV_NOISE:
            continue                                       # already folded into the frames
        if typ == EV_ROT:
            k = int(np.rint(val / math.pi))
            thp = val - k * math.pi
            if thp == 0.0:
                continue
            sg = 1.0 - 2.0 * ((masks[layer, 0] >> np.uint64(q0)) & one).astype(np.float64)
            c, s = math.cos(thp / 2), math.sin(thp / 2) * sg
            U = np.empty((T, 2, 2), dtype=np.complex128)
            U[:, 0, 0] = c
            U[:, 1, 1] = c
            U[:, 0, 1] = -1j * s
            U[:, 1, 0] = -1j * s
            F = _full_1q(U, pos[q0], nq)
            rho = F @ rho @ np.conj(np.transpose(F, (0, 2, 1)))
            continue
        if typ == EV_D1:
            sg = 1.0 - 2.0 * ((masks[layer, 1 + slot] >> np.uint64(q0)) & one).astype(np.float64)
            z = zbit[pos[q0]]
        else:
            sg = 1.0 - 2.0 * ((masks[layer, 3] >> np.uint64(slot)) & one).astype(np.float64)
            z = zbit[pos[q0]] if typ == EV_D2C else zbit[pos[q0]] * zbit[pos[q1]]
        ph = np.exp(-0.5j * val * sg[:, None] * z[None, :])
        rho = rho * ph[:, :, None] * np.conj(ph)[:, None, :]
    diag = np.real(np.einsum("tii->ti", rho))
    m = len(prog.measures)
    col = np.zeros(d, dtype=np.int64)
    for i, (b, _c) in enumerate(prog.measures):
        col |= ((idx >> pos[b]) & 1) << i
    flip = np.zeros(T, dtype=np.int64)
    for i, (b, _c) in enumerate(prog.measures):
        flip |= ((fx >> np.uint64(b)) & one).astype(np.int64) << i
    probs = np.zeros((T, 1 << m))
    for v in range(d):
        np.add.at(probs, (np.arange(T), col[v] ^ flip), diag[:, v])
    return probs
