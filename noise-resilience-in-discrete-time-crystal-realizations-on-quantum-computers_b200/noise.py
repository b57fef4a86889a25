"""Noise-model surface of the drop-in boundary (fast.py:76-86).

Mirrors the two qiskit-aer entry points the reference uses -- ``depolarizing_error(p, 1)`` and
``NoiseModel().add_all_qubit_quantum_error(error, ["u1","u2","u3"])`` -- and adapts a *real*
qiskit-aer ``NoiseModel`` through its ``to_dict()`` form.  Representable: single-qubit Pauli channels after
gates (all the simulator path of the reference uses) and single-qubit readout errors (the classical part of
device-calibrated noise, ``NoiseModel.from_backend``, fast.py:77-78 / SURVEY.md 8f-4); thermal-relaxation and other
non-Pauli channels raise ``ValueError`` -- there is no silent fallback.
"""


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
            self._all[nm] = self._all[nm].compose(error) if nm in self._all else error
            if nm not in self.basis_gates:
                self.basis_gates.append(nm)

    def add_quantum_error(self, error, instructions, qubits, warnings=True):
        if isinstance(instructions, str):
            instructions = [instructions]
        (q,) = tuple(qubits)
        for nm in instructions:
            key = (nm, int(q))
            self._local[key] = self._local[key].compose(error) if key in self._local else error

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
        """Pauli probabilities (px,py,pz) applied after gate `name` on `qubit`, or None."""
        e = self._local.get((name, qubit))
        if e is None:
            e = self._all.get(name)
        if e is None or e.is_ideal():
            return None
        return e.probs

    @property
    def noise_instructions(self):
        return sorted(set(self._all) | {k[0] for k in self._local})


_PAULI_NAMES = {"id": "I", "x": "X", "y": "Y", "z": "Z"}


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
        acc = {"I": 0.0, "X": 0.0, "Y": 0.0, "Z": 0.0}
        for circ, p in zip(err["instructions"], err["probabilities"]):
            label = "I"
            for inst in circ:
                name = inst["name"]
                if len(inst.get("qubits", [0])) != 1:
                    raise ValueError("multi-qubit noise instructions are not supported")
                if name in _PAULI_NAMES:
                    this = _PAULI_NAMES[name]
                elif name == "pauli":
                    this = str(inst["params"][0])
                    if this not in acc:
                        raise ValueError(f"unsupported Pauli string {this!r}")
                else:
                    raise ValueError(f"non-Pauli noise instruction {name!r} is not supported")
                # product of Paulis modulo phase: codes I=0,X=1,Y=2,Z=3 multiply by xor
                label = "IXYZ"["IXYZ".index(label) ^ "IXYZ".index(this)]
            acc[label] += float(p)
        e = PauliError(acc["X"], acc["Y"], acc["Z"])
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
