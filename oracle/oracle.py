"""CPU oracle: numpy restatement of the Aer path the reference scripts run.

TEST INFRASTRUCTURE ONLY.  Nothing under the product package may import this module; it is the
checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

PARITY STATUS: the arithmetic lives in qiskit-aer (third party, unpinned, absent from
/root/reference and from this image).  This restatement is therefore *pinned only statistically*
(3 sigma, chi^2/dof ~ 1) against the reference's committed 1024-shot CSVs and exactly against
the op multisets in its gate_counts_*.csv; at the 1e-10 / 1e-8 level parity is UNPINNED by the
reference and anchored on the published gate definitions below.

What is restated (reference call sites):
  * backend.run(circ, shots).result().get_counts()  fast.py:211-212  -> run_counts()
  * AerSimulator method "automatic": density_matrix when noise and shots > 2^n, otherwise one
    statevector noise trajectory per shot                         -> choose_method()
  * depolarizing_error(p,1) on ["u1","u2","u3"]      fast.py:85-86   -> PauliNoise
  * compute_z_expectation                            fast.py:92-109 -> compute_z_expectation()

Conventions: Qiskit little-endian (qubit k = bit k of the basis index); gate matrices per the
Qiskit circuit library: u3(t,p,l) = [[cos t/2, -e^{il} sin t/2],[e^{ip} sin t/2, e^{i(p+l)} cos t/2]],
rz(t) = diag(e^{-it/2}, e^{it/2}), rx/ry = exp(-i t X/2), exp(-i t Y/2), rzz = exp(-i t ZZ/2).
Execution is deliberately gate by gate, one pass over the state per op, like Aer without fusion.
"""
import math
import numpy as np

from . import philox

PAULI = {
    0: np.eye(2, dtype=np.complex128),
    1: np.array([[0, 1], [1, 0]], dtype=np.complex128),
    2: np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    3: np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


def u3_matrix(theta, phi, lam):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -np.exp(1j * lam) * s],
                     [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]], dtype=np.complex128)


def gate_matrix(name, params=()):
    """2x2 matrix of a named single-qubit gate (Qiskit circuit-library definitions)."""
    p = params
    if name in ("u3", "u"):
        return u3_matrix(p[0], p[1], p[2])
    if name == "u2":
        return u3_matrix(math.pi / 2, p[0], p[1])
    if name in ("u1", "p"):
        return np.array([[1, 0], [0, np.exp(1j * p[0])]], dtype=np.complex128)
    if name == "rz":
        return np.array([[np.exp(-0.5j * p[0]), 0], [0, np.exp(0.5j * p[0])]], dtype=np.complex128)
    if name == "rx":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if name == "ry":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if name == "h":
        return np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2)
    if name == "x":
        return PAULI[1].copy()
    if name == "y":
        return PAULI[2].copy()
    if name == "z":
        return PAULI[3].copy()
    if name == "id":
        return PAULI[0].copy()
    if name == "s":
        return np.diag([1, 1j]).astype(np.complex128)
    if name == "sdg":
        return np.diag([1, -1j]).astype(np.complex128)
    if name == "t":
        return np.diag([1, np.exp(0.25j * math.pi)]).astype(np.complex128)
    if name == "tdg":
        return np.diag([1, np.exp(-0.25j * math.pi)]).astype(np.complex128)
    if name == "sx":
        return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=np.complex128)
    if name == "sxdg":
        return 0.5 * np.array([[1 - 1j, 1 + 1j], [1 + 1j, 1 - 1j]], dtype=np.complex128)
    raise ValueError(f"oracle: unsupported gate {name}")


ONE_QUBIT = {"u3", "u", "u2", "u1", "p", "rz", "rx", "ry", "h", "x", "y", "z", "id", "s", "sdg",
             "t", "tdg", "sx", "sxdg"}


def _norm_op(op):
    if isinstance(op, tuple):
        name, qs, params, cs = op
    else:
        name, qs, params, cs = op.name, op.qubits, op.params, op.clbits
    return name, tuple(qs), tuple(params), tuple(cs)


# ----------------------------------------------------------------------------- state kernels
def apply_1q(psi, U, k):
    """psi: (..., 2^n) array; 2x2 U on qubit k (bit k of the last axis)."""
    shp = psi.shape
    v = psi.reshape(shp[:-1] + (-1, 2, 1 << k))
    out = np.empty_like(v)
    out[..., 0, :] = U[0, 0] * v[..., 0, :] + U[0, 1] * v[..., 1, :]
    out[..., 1, :] = U[1, 0] * v[..., 0, :] + U[1, 1] * v[..., 1, :]
    return out.reshape(shp)


def _bit(n, k):
    idx = np.arange(1 << n, dtype=np.int64)
    return (idx >> k) & 1


def apply_cx(psi, c, t):
    n = int(round(math.log2(psi.shape[-1])))
    idx = np.arange(1 << n, dtype=np.int64)
    src = np.where((idx >> c) & 1, idx ^ (1 << t), idx)
    return psi[..., src]


def apply_diag2(psi, a, b, d00, d01, d10, d11):
    """diagonal 2q gate, entries indexed (bit a, bit b)."""
    n = int(round(math.log2(psi.shape[-1])))
    ba, bb = _bit(n, a), _bit(n, b)
    tab = np.array([d00, d01, d10, d11], dtype=np.complex128)
    return psi * tab[2 * ba + bb]


def apply_op_unitary(psi, name, qs, params):
    if name in ONE_QUBIT:
        return apply_1q(psi, gate_matrix(name, params), qs[0])
    if name == "cx":
        return apply_cx(psi, qs[0], qs[1])
    if name == "cz":
        return apply_diag2(psi, qs[0], qs[1], 1, 1, 1, -1)
    if name == "rzz":
        e = np.exp(-0.5j * params[0])
        return apply_diag2(psi, qs[0], qs[1], e, np.conj(e), np.conj(e), e)
    if name == "swap":
        psi = apply_cx(psi, qs[0], qs[1])
        psi = apply_cx(psi, qs[1], qs[0])
        return apply_cx(psi, qs[0], qs[1])
    raise ValueError(f"oracle: unsupported op {name}")


# ----------------------------------------------------------------------------- noise model
class PauliNoise:
    """Single-qubit Pauli channel attached to gate names: {name: (pX, pY, pZ)}.

    depolarizing_error(p, 1) (fast.py:85): rho -> (1-p) rho + p I/2 == {I: 1-3p/4, X,Y,Z: p/4}.
    """

    def __init__(self, table=None, channels=None):
        self.table = dict(table or {})
        # general single-qubit channels as Kraus sets {name: [K_0, K_1, ...]} (2 x 2): the non-Pauli part of a device
        # noise model (fast.py:77-78); a gate name carries either a Pauli entry or a Kraus set; density-matrix method only
        self.channels = {nm: [np.asarray(k, dtype=np.complex128) for k in ks] for nm, ks in (channels or {}).items()}

    @classmethod
    def depolarizing(cls, p, names=("u1", "u2", "u3")):
        return cls({nm: (p / 4, p / 4, p / 4) for nm in names})

    @classmethod
    def thermal_relaxation(cls, t1, t2, time, excited_state_population=0.0, names=("u1", "u2", "u3")):
        return cls(channels={nm: thermal_relaxation_kraus(t1, t2, time, excited_state_population) for nm in names})

    def probs(self, name):
        return self.table.get(name)

    def kraus(self, name):
        return self.channels.get(name)

    def has_channels(self):
        return bool(self.channels)

    def is_ideal(self):
        return not any(any(v) for v in self.table.values()) and not self.channels


def thermal_relaxation_kraus(t1, t2, time, excited_state_population=0.0):
    """Kraus operators of qiskit-aer's thermal_relaxation_error, restated from its published construction: for T2 <= T1 the
    mixture {I: p_id, Z: p_z, reset to |0>: p_r0, reset to |1>: p_r1} with p_reset = 1 - exp(-time / T1),
    p_z = (1 - p_reset) (1 - exp(-time (1/T2 - 1/T1))) / 2, p_r0 = p_reset (1 - p1), p_r1 = p_reset p1; for T1 < T2 <= 2 T1 the
    Kraus operators obtained from the channel's Choi matrix
        [[1 - p1 p_reset, 0, 0, exp(-time / T2)], [0, p1 p_reset, 0, 0], [0, 0, p0 p_reset, 0], [exp(-time / T2), 0, 0, 1 - p0 p_reset]]."""
    p1 = float(excited_state_population)
    p0 = 1.0 - p1
    p_reset = 1.0 - np.exp(-time / t1)
    e2 = np.exp(-time / t2)
    if t2 <= t1:
        p_z = (1 - p_reset) * (1 - np.exp(-time * (1.0 / t2 - 1.0 / t1))) / 2
        p_r0, p_r1 = p_reset * p0, p_reset * p1
        p_id = 1 - p_z - p_r0 - p_r1
        I2, Z = np.eye(2), np.diag([1.0, -1.0])
        P00, P01 = np.array([[1.0, 0], [0, 0]]), np.array([[0, 1.0], [0, 0]])
        P10, P11 = np.array([[0, 0], [1.0, 0]]), np.array([[0, 0], [0, 1.0]])
        return [np.sqrt(p_id) * I2, np.sqrt(p_z) * Z, np.sqrt(p_r0) * P00, np.sqrt(p_r0) * P01,
                np.sqrt(p_r1) * P10, np.sqrt(p_r1) * P11]
    # Choi matrix C = sum_ij |i><j| (x) E(|i><j|), index (i, a), (j, b) with E(|i><j|)[a, b]
    choi = np.array([[1 - p1 * p_reset, 0, 0, e2], [0, p1 * p_reset, 0, 0], [0, 0, p0 * p_reset, 0],
                     [e2, 0, 0, 1 - p0 * p_reset]], dtype=np.complex128)
    w, v = np.linalg.eigh(choi)
    ks = []
    for lam, vec in zip(w, v.T):
        if lam > 1e-15:
            # vec index = 2 i + a  ->  K[a, i]
            ks.append(np.sqrt(lam) * vec.reshape(2, 2).T)
    return ks


def compact_ops(ops, n_qubits):
    """Drop idle qubits (Aer truncation): returns (ops', n_active, active_list)."""
    ops = [_norm_op(o) for o in ops]
    used = sorted({q for name, qs, _, _ in ops if name != "barrier" for q in qs})
    remap = {q: i for i, q in enumerate(used)}
    out = [(name, tuple(remap[q] for q in qs), params, cs) for name, qs, params, cs in ops
           if name != "barrier"]
    return out, len(used), used


def choose_method(n, shots, noise):
    """Aer 'automatic' (SURVEY A6): density_matrix iff noise present and shots > 2^n.  Non-Pauli channels always take the
    density matrix here (Aer would sample Kraus trajectories below that shot count; same outcome distribution)."""
    if noise is not None and getattr(noise, "has_channels", lambda: False)():
        return "density_matrix"
    if noise is not None and not noise.is_ideal() and shots > (1 << n):
        return "density_matrix"
    return "statevector"


# ----------------------------------------------------------------------------- runs
def run_statevector(ops, n, init=None):
    """Noiseless gate-by-gate statevector; measures are ignored. Returns psi (2^n,)."""
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0 if init is None else init] = 1.0
    for op in ops:
        name, qs, params, _ = _norm_op(op)
        if name in ("measure", "barrier"):
            continue
        psi = apply_op_unitary(psi, name, qs, params)
    return psi


def noise_sites(ops, noise):
    """Enumerate noisy gates in op order: list of (op_index, qubit, (pX,pY,pZ)). Site id = position."""
    sites = []
    for i, op in enumerate(ops):
        name, qs, _, _ = _norm_op(op)
        if noise is None or name not in ONE_QUBIT:
            continue
        pr = noise.probs(name)
        if pr is not None and any(pr):
            sites.append((i, qs[0], pr))
    return sites


def sample_paulis(ops, noise, seed, trajs):
    """Pauli codes [n_traj, n_sites] from the Philox contract (stream 0, index = site id)."""
    sites = noise_sites(ops, noise)
    trajs = np.asarray(trajs, dtype=np.uint64)
    codes = np.zeros((len(trajs), len(sites)), dtype=np.uint8)
    for s, (_, _, (px, py, pz)) in enumerate(sites):
        u = philox.uniform(seed, s, philox.STREAM_NOISE, trajs)
        codes[:, s] = philox.pauli_from_uniform(u, px, py, pz)
    return sites, codes


def run_trajectories(ops, n, noise, seed, trajs, init=None):
    """One statevector per trajectory id with sampled Pauli insertions after noisy gates.

    Returns psi [n_traj, 2^n].  This is Aer's statevector method under a Pauli noise model:
    each shot evolves a fresh |0..0> and draws one Pauli per noisy gate.
    """
    ops = [_norm_op(o) for o in ops]
    trajs = np.asarray(trajs, dtype=np.uint64)
    sites, codes = sample_paulis(ops, noise, seed, trajs)
    site_of_op = {i: s for s, (i, _, _) in enumerate(sites)}
    psi = np.zeros((len(trajs), 1 << n), dtype=np.complex128)
    psi[:, 0 if init is None else init] = 1.0
    for i, (name, qs, params, _) in enumerate(ops):
        if name in ("measure", "barrier"):
            continue
        psi = apply_op_unitary(psi, name, qs, params)
        if i in site_of_op:
            s = site_of_op[i]
            q = sites[s][1]
            for code in (1, 2, 3):
                sel = np.nonzero(codes[:, s] == code)[0]
                if len(sel):
                    psi[sel] = apply_1q(psi[sel], PAULI[code], q)
    return psi


def run_density_matrix(ops, n, noise, init=None):
    """Exact density matrix rho[row, col] (2^n x 2^n) with the Pauli channel after noisy gates."""
    ops = [_norm_op(o) for o in ops]
    d = 1 << n
    rho = np.zeros((d, d), dtype=np.complex128)
    i0 = 0 if init is None else init
    rho[i0, i0] = 1.0

    def conj_unitary(rho, name, qs, params):
        # rho -> U rho U^dagger : apply U on the row index, conj(U) on the column index
        rho = apply_op_unitary(rho.T, name, qs, params).T           # rows (axis 0 made last)
        rho = np.conj(apply_op_unitary(np.conj(rho), name, qs, params))  # columns
        return rho

    for name, qs, params, _ in ops:
        if name in ("measure", "barrier"):
            continue
        rho = conj_unitary(rho, name, qs, params)
        ks = noise.kraus(name) if (noise is not None and name in ONE_QUBIT and hasattr(noise, "kraus")) else None
        if ks is not None:
            # rho -> sum_k K rho K^dagger on the gate's qubit: K on the row index, conj(K) on the column index
            q = qs[0]
            acc = np.zeros_like(rho)
            for K in ks:
                r1 = apply_1q(rho.T, K, q).T
                acc = acc + np.conj(apply_1q(np.conj(r1), K, q))
            rho = acc
            continue
        pr = noise.probs(name) if (noise is not None and name in ONE_QUBIT) else None
        if pr is not None and any(pr):
            q = qs[0]
            px, py, pz = pr
            acc = (1.0 - px - py - pz) * rho
            for code, pp in ((1, px), (2, py), (3, pz)):
                if pp:
                    nm = {1: "x", 2: "y", 3: "z"}[code]
                    acc = acc + pp * conj_unitary(rho, nm, (q,), ())
            rho = acc
    return rho


def measured_map(ops):
    """[(qubit, clbit)] in op order; a clbit written twice keeps the last writer."""
    m = {}
    for op in ops:
        name, qs, _, cs = _norm_op(op)
        if name == "measure":
            m[cs[0]] = qs[0]
    return sorted((q, c) for c, q in m.items())


def outcome_probabilities(p_full, n, meas, n_clbits):
    """Marginalise |psi|^2 (or diag rho) [..., 2^n] onto classical-register values [..., 2^n_clbits]."""
    idx = np.arange(1 << n, dtype=np.int64)
    cval = np.zeros(1 << n, dtype=np.int64)
    for q, c in meas:
        cval |= ((idx >> q) & 1) << c
    out = np.zeros(p_full.shape[:-1] + (1 << n_clbits,), dtype=np.float64)
    flat = p_full.reshape(-1, 1 << n)
    oflat = out.reshape(-1, 1 << n_clbits)
    for r in range(flat.shape[0]):
        oflat[r] = np.bincount(cval, weights=flat[r], minlength=1 << n_clbits)
    return out


def sample_outcome(cum_probs, u):
    """Inverse-CDF sample: first index with cum > u (clamped)."""
    k = np.searchsorted(cum_probs, u, side="right")
    return np.minimum(k, len(cum_probs) - 1)


def counts_dict(values, n_clbits):
    """Aer get_counts format: binary keys, clbit 0 rightmost, zero-count keys omitted."""
    c = {}
    for v in values:
        key = format(int(v), f"0{n_clbits}b")
        c[key] = c.get(key, 0) + 1
    return c


def apply_readout(vals, readout, seed, shot_ids):
    """Classical readout errors (qiskit-aer ReadoutError semantics: readout[c][i][j] = P(recorded j | true i) for classical
    bit c): bit c of shot s flips when u = philox.uniform(seed, c, STREAM_READOUT, s) falls below its flip probability."""
    vals = np.array(vals, dtype=np.int64, copy=True)
    for c, m in (readout or {}).items():
        u = philox.uniform(seed, c, philox.STREAM_READOUT, np.asarray(shot_ids, dtype=np.uint64))
        bit = (vals >> c) & 1
        flip = np.where(bit == 0, u < m[0][1], u < m[1][0])
        vals ^= flip.astype(np.int64) << c
    return vals


def run_counts(ops, n_qubits, n_clbits, shots=1024, noise=None, seed=1234, method="automatic", readout=None):
    """Full restatement of backend.run(circ, shots).result().get_counts() (fast.py:211-212).

    Returns (counts, info) with info = {method, probabilities (DM / noiseless) or per-trajectory p}.
    Sampling contract: shot s draws u = philox.uniform(seed, 0 | s, STREAM_MEASURE, traj) where
    traj = s for trajectory runs and the index = s, traj = 0 for single-state runs.
    """
    ops_c, n, _ = compact_ops(ops, n_qubits)
    meas = measured_map(ops_c)
    if method == "automatic":
        method = choose_method(n, shots, noise)
    noisy = noise is not None and not noise.is_ideal()
    if method == "density_matrix":
        rho = run_density_matrix(ops_c, n, noise if noisy else None)
        probs = outcome_probabilities(np.real(np.diag(rho)).copy(), n, meas, n_clbits)
        u = philox.uniform(seed, np.arange(shots), philox.STREAM_MEASURE, 0)
        vals = apply_readout(sample_outcome(np.cumsum(probs), u), readout, seed, np.arange(shots))
        return counts_dict(vals, n_clbits), {"method": method, "probabilities": probs}
    if not noisy:
        psi = run_statevector(ops_c, n)
        probs = outcome_probabilities(np.abs(psi) ** 2, n, meas, n_clbits)
        u = philox.uniform(seed, np.arange(shots), philox.STREAM_MEASURE, 0)
        vals = apply_readout(sample_outcome(np.cumsum(probs), u), readout, seed, np.arange(shots))
        return counts_dict(vals, n_clbits), {"method": method, "probabilities": probs}
    trajs = np.arange(shots, dtype=np.uint64)
    vals = np.zeros(shots, dtype=np.int64)
    ptraj = np.zeros((shots, 1 << n_clbits))
    chunk = max(1, (1 << 22) >> n)
    for a in range(0, shots, chunk):
        tr = trajs[a:a + chunk]
        psi = run_trajectories(ops_c, n, noise, seed, tr)
        probs = outcome_probabilities(np.abs(psi) ** 2, n, meas, n_clbits)
        ptraj[a:a + chunk] = probs
        u = philox.uniform(seed, 0, philox.STREAM_MEASURE, tr)
        for r in range(len(tr)):
            vals[a + r] = sample_outcome(np.cumsum(probs[r]), u[r])
    vals = apply_readout(vals, readout, seed, np.arange(shots))
    return counts_dict(vals, n_clbits), {"method": method, "trajectory_probabilities": ptraj}


def compute_z_expectation(counts, num_qubits):
    """fast.py:92-109 restated: <Z_k> = (N0 - N1)/shots with bit k = bitstring[::-1][k]."""
    total = sum(counts.values())
    out = []
    for k in range(num_qubits):
        p0 = sum(c for b, c in counts.items() if b[::-1][k] == "0")
        out.append((p0 - (total - p0)) / total)
    return out


# ----------------------------------------------------------------------------- light-cone exact values
def lightcone_zq(L, g, hs, phis, t, q, p, echo=False, polarization="x", period_fn=None):
    """Exact noisy <Z_q> after t periods (forward) or t forward + t inverse periods (echo) for a
    vacuum start, computed on the sub-chain [q-r, q+r], r = n_periods-1.

    RZ/RZZ commute with Z_q, so the support of the Heisenberg-evolved Z_q grows by one site per
    period only through the kick layer; bonds leaving the sub-chain drop out exactly.  The Hadamard
    test signal of fast.py:125-147 is (1-p)^6 * <Z_q> (six depolarized ancilla u2 gates).
    period_fn(step) -> list of high-level gates on sites 1..L (defaults to dtc_circuits.uf_gates).
    """
    from . import dtc_circuits as C
    n_periods = 2 * t if echo else t
    if n_periods == 0:
        return 1.0
    r = n_periods - 1
    lo, hi = max(0, q - r), min(L - 1, q + r)
    sites = list(range(lo, hi + 1))
    m = len(sites)
    pos = {s: i for i, s in enumerate(sites)}

    def restrict(gates):
        out = []
        for name, qs, params, cs in gates:
            ss = [x - 1 for x in qs]                      # uf_gates uses circuit qubits 1..L
            if all(s in pos for s in ss):
                out.append((name, tuple(pos[s] for s in ss), params, cs))
        return out

    def period(step):
        if period_fn is not None:
            return period_fn(step)
        return C.uf_gates(L, g, phis, hs, polarization, time_step=step)

    ops = []
    for step in range(t):
        ops.extend(restrict(period(step)))
    if echo:
        for step in range(t - 1, -1, -1):
            ops.extend(restrict(C.inverse_gates(period(step))))
    low = C.lower_level0(ops)
    noise = PauliNoise.depolarizing(p) if p else None
    rho = run_density_matrix(low, m, noise)
    diag = np.real(np.diag(rho))
    z = 1.0 - 2.0 * _bit(m, pos[q])
    return float(np.dot(diag, z))
