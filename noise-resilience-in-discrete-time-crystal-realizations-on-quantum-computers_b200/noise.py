"""Noise-model surface of the drop-in boundary (fast.py:76-86).

Mirrors the two qiskit-aer entry points the reference uses -- ``depolarizing_error(p, 1)`` and
``NoiseModel().add_all_qubit_quantum_error(error, ["u1","u2","u3"])`` -- and adapts a *real*
qiskit-aer ``NoiseModel`` through its ``to_dict()`` form.  Representable: single-qubit Pauli channels after
gates (all the simulator path of the reference uses), single-qubit readout errors (the classical part of
device-calibrated noise, ``NoiseModel.from_backend``, fast.py:77-78 / SURVEY.md 8f-4) and general single-qubit channels
(``ChannelError``: thermal relaxation, amplitude / phase damping, Kraus sets and ``reset`` instructions of a device model).
Pauli channels run on every method; a non-Pauli channel is executed exactly by the density-matrix method (n <= 13) and is
refused with ``ValueError`` on the trajectory path (norm-dependent Kraus sampling does not fit the Pauli-frame engine) --
there is no silent fallback.  Errors on multi-qubit gates are not supported.
"""
import math

import numpy as np


class PauliError:
    """Single-qubit Pauli mixture {I: 1-px-py-pz, X: px, Y: py, Z: pz}."""

    def __init__(self, px, py, pz):
        for v in (px, py, pz):
            if v < 0:
                raise ValueError("negative Pauli probability")
        if px + py + pz > 1 + 1e-12:
            raise ValueError("Pauli probabilities exceed 1")
        self.px, self.py, self.pz = float(px), float(py), float(pz)

    @property
    def probs(self):
        return (self.px, self.py, self.pz)

    def compose(self, other):
        """Channel composition other o self (both Pauli, so the result is Pauli)."""
        a = (1 - self.px - self.py - self.pz, self.px, self.py, self.pz)
        b = (1 - other.px - other.py - other.pz, other.px, other.py, other.pz)
        # Pauli product modulo phase: codes I=0,X=1,Y=2,Z=3 multiply by xor
        out = [0.0] * 4
        for i in range(4):
            for j in range(4):
                out[i ^ j] += a[i] * b[j]
        return PauliError(out[1], out[2], out[3])

    def is_ideal(self):
        return self.px == 0 and self.py == 0 and self.pz == 0

    def superop(self):
        return ChannelError.from_kraus([math.sqrt(max(0.0, 1 - self.px - self.py - self.pz)) * _PAULI_MATS["I"],
                                        math.sqrt(self.px) * _PAULI_MATS["X"], math.sqrt(self.py) * _PAULI_MATS["Y"],
                                        math.sqrt(self.pz) * _PAULI_MATS["Z"]]).S


_PAULI_MATS = {"I": np.eye(2, dtype=complex), "X": np.array([[0, 1], [1, 0]], dtype=complex),
               "Y": np.array([[0, -1j], [1j, 0]], dtype=complex), "Z": np.diag([1.0 + 0j, -1.0])}


class ChannelError:
    """General single-qubit channel as a 4 x 4 superoperator S on vec(rho) with index = row + 2 col (the layout of a
    (row bit, column bit) block of the density-matrix engine): out[r' + 2c'] = sum S[r' + 2c', r + 2c] rho[r, c];
    for Kraus operators K_k:  S = sum_k conj(K_k) (x) K_k."""

    def __init__(self, superop):
        S = np.asarray(superop, dtype=np.complex128)
        if S.shape != (4, 4):
            raise ValueError("a single-qubit channel is a 4 x 4 superoperator")
        # trace preservation: sum over r' of S[(r', r'), (r, c)] = delta(r, c)
        tp = S[0] + S[3]
        if np.abs(tp - np.array([1, 0, 0, 1])).max() > 1e-9:
            raise ValueError("channel is not trace preserving")
        self.S = S

    @classmethod
    def from_kraus(cls, kraus):
        S = np.zeros((4, 4), dtype=np.complex128)
        for K in kraus:
            K = np.asarray(K, dtype=np.complex128)
            if K.shape != (2, 2):
                raise ValueError("only single-qubit (2 x 2) Kraus operators are supported")
            S += np.kron(np.conj(K), K)
        return cls(S)

    @property
    def probs(self):
        """Hashable identity (used as a cache key beside PauliError.probs)."""
        return tuple(np.round(self.S.reshape(-1), 15).tolist())

    def compose(self, other):
        """Channel composition other o self."""
        So = other.S if isinstance(other, ChannelError) else other.superop()
        return ChannelError(So @ self.S)

    def is_ideal(self):
        return np.abs(self.S - np.eye(4)).max() == 0.0

    def _ptm(self):
        """Pauli-transfer matrix R_ab = tr(P_a E(P_b)) / 2 (complex in general)."""
        R = np.zeros((4, 4), dtype=np.complex128)
        labels = "IXYZ"
        for b, lb in enumerate(labels):
            out = (self.S @ _PAULI_MATS[lb].reshape(-1, order="F")).reshape(2, 2, order="F")
            for a, la in enumerate(labels):
                R[a, b] = np.trace(_PAULI_MATS[la] @ out) / 2
        return R

    def pauli_twirl(self):
        """The Pauli channel obtained by twirling this channel over the Pauli group (its Pauli-transfer matrix restricted to
        the diagonal): an APPROXIMATION that keeps the decay rates of <X>, <Y>, <Z> and drops the non-unital drift (e.g. the
        relaxation towards |0>).  Opt-in only (NoiseModel.pauli_twirled): it is what lets a device-like noise model run on the
        trajectory engine at n > 13; exact results need the density-matrix method."""
        R = self._ptm().real
        lx, ly, lz = R[1, 1], R[2, 2], R[3, 3]
        px, py, pz = (1 + lx - ly - lz) / 4, (1 - lx + ly - lz) / 4, (1 - lx - ly + lz) / 4
        if min(px, py, pz) < -1e-12:
            raise ValueError("the Pauli twirl of this channel is not a probability mixture")
        return PauliError(max(px, 0.0), max(py, 0.0), max(pz, 0.0))

    def as_pauli(self, tol=1e-13):
        """The PauliError this channel equals, or None: a Pauli mixture has a diagonal Pauli-transfer matrix with
        non-negative mixture weights."""
        # Pauli-transfer entries R_ab = tr(P_a E(P_b)) / 2
        R = np.zeros((4, 4))
        labels = "IXYZ"
        for b, lb in enumerate(labels):
            v = _PAULI_MATS[lb].reshape(-1, order="F")          # vec with index = row + 2 col
            out = (self.S @ v).reshape(2, 2, order="F")
            for a, la in enumerate(labels):
                val = np.trace(_PAULI_MATS[la] @ out) / 2
                if abs(val.imag) > tol:
                    return None
                R[a, b] = val.real
        if np.abs(R - np.diag(np.diag(R))).max() > tol:
            return None
        lx, ly, lz = R[1, 1], R[2, 2], R[3, 3]
        px, py, pz = (1 + lx - ly - lz) / 4, (1 - lx + ly - lz) / 4, (1 - lx - ly + lz) / 4
        if min(px, py, pz) < -tol or px + py + pz > 1 + tol:
            return None
        return PauliError(max(px, 0.0), max(py, 0.0), max(pz, 0.0))


def kraus_error(kraus):
    """qiskit_aer.noise.kraus_error for single-qubit Kraus sets."""
    e = ChannelError.from_kraus(kraus)
    return e.as_pauli() or e


def amplitude_damping_error(param_amp, excited_state_population=0.0):
    """qiskit_aer.noise.amplitude_damping_error: decay 1 -> 0 with probability gamma towards a thermal state with
    excited-state population p1 (generalised amplitude damping)."""
    g, p1 = float(param_amp), float(excited_state_population)
    if not (0 <= g <= 1 and 0 <= p1 <= 1):
        raise ValueError("amplitude damping parameters out of range")
    s = math.sqrt(1 - g)
    ks = [math.sqrt(1 - p1) * np.array([[1, 0], [0, s]]), math.sqrt(1 - p1) * np.array([[0, math.sqrt(g)], [0, 0]]),
          math.sqrt(p1) * np.array([[s, 0], [0, 1]]), math.sqrt(p1) * np.array([[0, 0], [math.sqrt(g), 0]])]
    return ChannelError.from_kraus(ks)


def phase_damping_error(param_phase):
    """qiskit_aer.noise.phase_damping_error: off-diagonal elements shrink by sqrt(1 - lambda) (a Pauli-Z mixture)."""
    lam = float(param_phase)
    if not 0 <= lam <= 1:
        raise ValueError("phase damping parameter out of range")
    return PauliError(0.0, 0.0, (1 - math.sqrt(1 - lam)) / 2)


def thermal_relaxation_error(t1, t2, time, excited_state_population=0.0):
    """qiskit_aer.noise.thermal_relaxation_error: populations relax towards (1 - p1, p1) with probability
    p_reset = 1 - exp(-time / T1), coherences decay by exp(-time / T2) (T2 <= 2 T1).  Aer builds it as a mixture of
    {I, Z, reset to 0, reset to 1} for T2 <= T1 and from its Choi matrix otherwise; both are this superoperator."""
    t1, t2, time, p1 = float(t1), float(t2), float(time), float(excited_state_population)
    if time < 0 or t1 <= 0 or t2 <= 0:
        raise ValueError("thermal relaxation needs T1 > 0, T2 > 0, time >= 0")
    if t2 - 2 * t1 > 0:
        raise ValueError("thermal relaxation needs T2 <= 2 T1")
    if not 0 <= p1 <= 1:
        raise ValueError("excited-state population out of range")
    p_reset = 1.0 - (math.exp(-time / t1) if math.isfinite(t1) else 1.0)
    e2 = math.exp(-time / t2) if math.isfinite(t2) else 1.0
    p0 = 1.0 - p1
    S = np.zeros((4, 4), dtype=np.complex128)
    # index = row + 2 col: 0 = rho00, 1 = rho10, 2 = rho01, 3 = rho11
    S[0, 0] = 1 - p_reset + p_reset * p0
    S[0, 3] = p_reset * p0
    S[3, 3] = 1 - p_reset + p_reset * p1
    S[3, 0] = p_reset * p1
    S[1, 1] = S[2, 2] = e2
    e = ChannelError(S)
    return e.as_pauli() or e


class ReadoutError:
    """qiskit_aer.noise.ReadoutError for one qubit: probabilities[i][j] = P(recorded j | true outcome i)."""

    def __init__(self, probabilities):
        m = [[float(x) for x in row] for row in probabilities]
        if len(m) != 2 or any(len(r) != 2 for r in m):
            raise ValueError("only single-qubit readout errors (2 x 2 assignment matrices) are supported")
        for r in m:
            if min(r) < 0 or abs(sum(r) - 1.0) > 1e-9:
                raise ValueError("rows of a readout assignment matrix must be probability vectors")
        self.probabilities = m

    def is_ideal(self):
        return self.probabilities[0][1] == 0.0 and self.probabilities[1][0] == 0.0


def depolarizing_error(param, num_qubits=1):
    """qiskit_aer.noise.depolarizing_error: rho -> (1-p) rho + p I/2 (fast.py:85)."""
    if num_qubits != 1:
        raise ValueError("only single-qubit depolarizing errors are supported on this path")
    if not 0 <= param <= 4.0 / 3.0:
        raise ValueError("depolarizing parameter out of range")
    return PauliError(param / 4, param / 4, param / 4)


def pauli_error(terms):
    """qiskit_aer.noise.pauli_error([('X', p), ...]) for single-qubit labels."""
    acc = {"I": 0.0, "X": 0.0, "Y": 0.0, "Z": 0.0}
    for label, p in terms:
        if label not in acc:
            raise ValueError(f"unsupported Pauli label {label!r}")
        acc[label] += p
    return PauliError(acc["X"], acc["Y"], acc["Z"])


class NoiseModel:
    """Subset of qiskit_aer.noise.NoiseModel used by the reference."""

    def __init__(self, basis_gates=None):
        self._all = {}        # gate name -> PauliError
        self._local = {}      # (gate name, qubit) -> PauliError
        self._ro_all = None   # ReadoutError on every measured qubit
        self._ro_local = {}   # qubit -> ReadoutError
        self.basis_gates = list(basis_gates or ["cx", "id", "rz", "sx"])

    def add_all_qubit_quantum_error(self, error, instructions, warnings=True):
        if isinstance(instructions, str):
            instructions = [instructions]
        for nm in instructions:
            # adding to an instruction that already has an error composes them
            self._all[nm] = _compose(self._all[nm], error) if nm in self._all else error
            if nm not in self.basis_gates:
                self.basis_gates.append(nm)

    def add_quantum_error(self, error, instructions, qubits, warnings=True):
        if isinstance(instructions, str):
            instructions = [instructions]
        (q,) = tuple(qubits)
        for nm in instructions:
            key = (nm, int(q))
            self._local[key] = _compose(self._local[key], error) if key in self._local else error

    def add_all_qubit_readout_error(self, error, warnings=True):
        self._ro_all = error if isinstance(error, ReadoutError) else ReadoutError(error)

    def add_readout_error(self, error, qubits, warnings=True):
        (q,) = tuple(qubits)
        self._ro_local[int(q)] = error if isinstance(error, ReadoutError) else ReadoutError(error)

    def has_gate_noise(self):
        return not (all(e.is_ideal() for e in self._all.values()) and all(e.is_ideal() for e in self._local.values()))

    def has_readout_noise(self):
        return (self._ro_all is not None and not self._ro_all.is_ideal()) or \
            any(not e.is_ideal() for e in self._ro_local.values())

    def is_ideal(self):
        return not self.has_gate_noise() and not self.has_readout_noise()

    def lookup_readout(self, qubit):
        """2 x 2 assignment matrix P(recorded | true) of the measurement of `qubit`, or None."""
        e = self._ro_local.get(int(qubit), self._ro_all)
        if e is None or e.is_ideal():
            return None
        return e.probabilities

    def lookup(self, name, qubit):
        """Error applied after gate `name` on `qubit`: Pauli probabilities (px, py, pz), the 4 x 4 superoperator of a
        non-Pauli channel (numpy array, see ChannelError), or None."""
        e = self._local.get((name, qubit))
        if e is None:
            e = self._all.get(name)
        if e is None or e.is_ideal():
            return None
        return e.S if isinstance(e, ChannelError) else e.probs

    def has_channel_noise(self):
        """True if some error is not a Pauli mixture (density-matrix method only)."""
        return any(isinstance(e, ChannelError) and not e.is_ideal()
                   for e in list(self._all.values()) + list(self._local.values()))

    def pauli_twirled(self):
        """Copy of this model with every non-Pauli channel replaced by its Pauli twirl (ChannelError.pauli_twirl): an explicit,
        opt-in approximation for running device-like noise on the trajectory engine; readout errors are kept."""
        out = NoiseModel(self.basis_gates)
        out._all = {k: (e.pauli_twirl() if isinstance(e, ChannelError) else e) for k, e in self._all.items()}
        out._local = {k: (e.pauli_twirl() if isinstance(e, ChannelError) else e) for k, e in self._local.items()}
        out._ro_all, out._ro_local = self._ro_all, dict(self._ro_local)
        return out

    @property
    def noise_instructions(self):
        return sorted(set(self._all) | {k[0] for k in self._local})


def _compose(first, then):
    """Channel composition `then` o `first` of PauliError / ChannelError objects (Pauli o Pauli stays Pauli)."""
    if isinstance(first, PauliError) and isinstance(then, PauliError):
        return first.compose(then)
    a = first if isinstance(first, ChannelError) else ChannelError(first.superop())
    out = a.compose(then)
    return out.as_pauli() or out


_PAULI_NAMES = {"id": "I", "x": "X", "y": "Y", "z": "Z"}
_RESET_KRAUS = [np.array([[1, 0], [0, 0]], dtype=complex), np.array([[0, 1], [0, 0]], dtype=complex)]


def _complex_matrix(m):
    """A 2 x 2 matrix from a numpy array, nested complex lists or qiskit's serialised [[re, im], ...] form."""
    a = np.asarray(m)
    if a.dtype != object and a.ndim == 3 and a.shape[-1] == 2 and not np.iscomplexobj(a):
        a = a[..., 0] + 1j * a[..., 1]
    a = np.asarray(a, dtype=np.complex128)
    if a.shape != (2, 2):
        raise ValueError("only single-qubit (2 x 2) noise operators are supported")
    return a


def _instruction_superop(inst):
    """Superoperator of one noise-circuit instruction of NoiseModel.to_dict(): Pauli gates, 'pauli', 'reset', 'kraus',
    'unitary'."""
    name = inst["name"]
    if len(inst.get("qubits", [0])) != 1:
        raise ValueError("multi-qubit noise instructions are not supported")
    if name in _PAULI_NAMES:
        return ChannelError.from_kraus([_PAULI_MATS[_PAULI_NAMES[name]]]).S
    if name == "pauli":
        label = str(inst["params"][0])
        if label not in _PAULI_MATS:
            raise ValueError(f"unsupported Pauli string {label!r}")
        return ChannelError.from_kraus([_PAULI_MATS[label]]).S
    if name == "reset":
        return ChannelError.from_kraus(_RESET_KRAUS).S
    if name == "kraus":
        return ChannelError.from_kraus([_complex_matrix(k) for k in inst["params"]]).S
    if name == "unitary":
        return ChannelError.from_kraus([_complex_matrix(inst["params"][0])]).S
    raise ValueError(f"noise instruction {name!r} is not supported")


def _from_dict(d):
    nm = NoiseModel()
    for err in d.get("errors", []):
        if err.get("type", "qerror") == "roerror":
            ro = ReadoutError(err["probabilities"])
            gate_qubits = err.get("gate_qubits")
            if gate_qubits:
                for gq in gate_qubits:
                    nm.add_readout_error(ro, gq)
            else:
                nm.add_all_qubit_readout_error(ro)
            continue
        if err.get("type", "qerror") != "qerror":
            raise ValueError(f"unsupported noise entry type {err.get('type')!r}")
        pauli_only = all(inst["name"] in _PAULI_NAMES or inst["name"] == "pauli"
                         for circ in err["instructions"] for inst in circ)
        if pauli_only:
            # Pauli mixture: probabilities accumulate exactly (no round trip through a superoperator)
            acc = {"I": 0.0, "X": 0.0, "Y": 0.0, "Z": 0.0}
            for circ, p in zip(err["instructions"], err["probabilities"]):
                label = "I"
                for inst in circ:
                    if len(inst.get("qubits", [0])) != 1:
                        raise ValueError("multi-qubit noise instructions are not supported")
                    this = _PAULI_NAMES[inst["name"]] if inst["name"] in _PAULI_NAMES else str(inst["params"][0])
                    if this not in acc:
                        raise ValueError(f"unsupported Pauli string {this!r}")
                    # product of Paulis modulo phase: codes I=0,X=1,Y=2,Z=3 multiply by xor
                    label = "IXYZ"["IXYZ".index(label) ^ "IXYZ".index(this)]
                acc[label] += float(p)
            e = PauliError(acc["X"], acc["Y"], acc["Z"])
        else:
            # mixture of noise circuits: sum_k p_k (product of the superoperators of circuit k's instructions)
            S = np.zeros((4, 4), dtype=np.complex128)
            for circ, p in zip(err["instructions"], err["probabilities"]):
                Sk = np.eye(4, dtype=np.complex128)
                for inst in circ:
                    Sk = _instruction_superop(inst) @ Sk
                S += float(p) * Sk
            ch = ChannelError(S)
            e = ch.as_pauli() or ch
        gate_qubits = err.get("gate_qubits")
        if gate_qubits:
            for gq in gate_qubits:
                nm.add_quantum_error(e, err["operations"], gq)
        else:
            nm.add_all_qubit_quantum_error(e, err["operations"])
    return nm


def as_noise_model(obj):
    """None | native NoiseModel | qiskit-aer NoiseModel (via to_dict) | dict -> NoiseModel or None."""
    if obj is None:
        return None
    if isinstance(obj, NoiseModel):
        return None if obj.is_ideal() else obj
    if isinstance(obj, dict):
        nm = _from_dict(obj)
    elif hasattr(obj, "to_dict"):
        nm = _from_dict(obj.to_dict())
    else:
        raise TypeError(f"cannot interpret {type(obj).__name__} as a noise model")
    return None if nm.is_ideal() else nm
