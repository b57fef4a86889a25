"""Estimator path of the reference's energy scripts on top of the drop-in backend.

    estimator = BackendEstimatorV2(backend=backend)                                   # energy.py:166
    results = estimator.run([(circ_tnoise, hamiltonian)]).result()                    # energy.py:168
    expval = results[0].data.evs                                                      # energy.py:169

With qiskit installed, qiskit's own `BackendEstimatorV2` drives `DTCSimulator.run()` (it only needs `run(circuits,
shots=...)` and `get_counts()`); this module is the same algorithm without qiskit, so the path exists in this image:
the observable's Pauli terms are grouped into qubit-wise commuting sets, every set becomes ONE measurement circuit
(the state-preparation circuit + a basis change H / Sdg-H on the measured qubits + measure), all measurement circuits
go through one pipelined `backend.run(list, shots)` call, and `evs = sum_k c_k <P_k>` comes from the parities of the
counts.  Shots per circuit = ceil(1 / precision^2); the default precision 0.015625 gives 4096 (qiskit's default).

Pauli labels follow qiskit: the RIGHTMOST character acts on qubit 0.  The reference's `get_hamiltonian`
(energy.py:83-102) writes the 'Z' of `hs[i]` at string position i, i.e. on qubit L-1-i, while its circuit puts
RZ(hs[i]) on qubit i (energy.py:131-132): `dtc_hamiltonian` reproduces those labels verbatim, so whoever builds H the
reference's way gets the reference's numbers, quirk included.
"""
import math

import numpy as np

from .ir import QuantumCircuit, as_circuit


def dtc_hamiltonian(L, g, phis, hs):
    """The reference's get_hamiltonian (energy.py:83-102) as a list of (label, coefficient): sum_i hs[i] Z + sum_i phis[i] ZZ
    + pi g sum_i X, each written at STRING position i of an L-character label (see the module note on label order)."""
    ham = []
    ident = "I" * L
    for i in range(L):
        ham.append((ident[:i] + "Z" + ident[i + 1:], float(hs[i])))
    for i in range(L - 1):
        ham.append((ident[:i] + "ZZ" + ident[i + 2:], float(phis[i])))
    for i in range(L):
        ham.append((ident[:i] + "X" + ident[i + 1:], float(g) * math.pi))
    return ham


def observable_terms(obs, num_qubits=None):
    """[(label, coeff)] from a SparsePauliOp-like object (`to_list()`), a list of pairs, a {label: coeff} dict or a label."""
    if isinstance(obs, str):
        terms = [(obs, 1.0)]
    elif isinstance(obs, dict):
        terms = list(obs.items())
    elif hasattr(obs, "to_list"):
        terms = list(obs.to_list())
    else:
        terms = list(obs)
    out = []
    for label, c in terms:
        label = str(label)
        if any(ch not in "IXYZ" for ch in label):
            raise ValueError(f"unsupported Pauli label {label!r}")
        if num_qubits is not None and len(label) != num_qubits:
            raise ValueError(f"Pauli label {label!r} does not match the circuit's {num_qubits} qubits")
        c = complex(c)
        if abs(c.imag) > 1e-12 * max(1.0, abs(c.real)):
            raise ValueError("observable coefficients must be real")
        out.append((label, c.real))
    return out


def group_qubitwise_commuting(terms):
    """Greedy grouping: a term joins the first group whose measurement basis it does not contradict.  Returns a list of
    (basis, members): basis = per-qubit Pauli character ('I' where no member acts), qubit 0 FIRST; members = term indices."""
    groups = []
    for k, (label, _c) in enumerate(terms):
        word = label[::-1]                                 # word[q] = Pauli on qubit q
        for basis, members in groups:
            if all(a == "I" or b == "I" or a == b for a, b in zip(word, basis)):
                for q, ch in enumerate(word):
                    if ch != "I":
                        basis[q] = ch
                members.append(k)
                break
        else:
            groups.append((list(word), [k]))
    return [("".join(b), m) for b, m in groups]


def measurement_circuit(circuit, basis):
    """circuit + basis change + measurement of the qubits where `basis` (qubit 0 first) is not 'I'; clbit i <- i-th such qubit."""
    circ = as_circuit(circuit)
    if any(op.name == "measure" for op in circ.ops):
        raise ValueError("the estimator adds its own measurements: pass a circuit without measure instructions")
    qubits = [q for q, ch in enumerate(basis) if ch != "I"]
    out = QuantumCircuit(circ.num_qubits, max(len(qubits), 1), circ.name)
    out.ops = list(circ.ops)
    out.global_phase = circ.global_phase
    for q in qubits:
        if basis[q] == "X":
            out.h(q)
        elif basis[q] == "Y":
            out.sdg(q)
            out.h(q)
    for i, q in enumerate(qubits):
        out.measure(q, i)
    return out, qubits


def pauli_expectation(counts, positions):
    """<P> of a Pauli word from the counts of its group's circuit: mean parity of the clbits `positions`."""
    tot, acc = 0, 0
    for key, n in counts.items():
        bits = key.replace(" ", "")[::-1]
        par = 0
        for p in positions:
            par ^= bits[p] == "1"
        acc += -n if par else n
        tot += n
    return acc / tot


class _DataBin:
    def __init__(self, evs, stds, shots):
        self.evs, self.stds, self.shots = evs, stds, shots


class PubResult:
    def __init__(self, data, metadata):
        self.data, self.metadata = data, metadata


class EstimatorJob:
    def __init__(self, results):
        self._results = results

    def result(self):
        return self._results


class BackendEstimatorV2:
    """Same call surface as qiskit.primitives.BackendEstimatorV2 for what energy.py uses: run([(circuit, observable)]).
    `circuit` must act on the physical qubits the observable's labels refer to (as the reference's transpiled circuit
    does); circuits wider than the labels are accepted when the extra qubits are idle."""

    def __init__(self, backend, options=None):
        self.backend = backend
        opts = dict(options or {})
        self.default_precision = float(opts.get("default_precision", 0.015625))
        self.abelian_grouping = bool(opts.get("abelian_grouping", True))
        self.seed_simulator = opts.get("seed_simulator")

    def run(self, pubs, precision=None):
        results = []
        for pub in pubs:
            if not isinstance(pub, (tuple, list)) or len(pub) < 2:
                raise ValueError("a pub is (circuit, observable[, parameter_values[, precision]])")
            if len(pub) > 2 and pub[2] is not None and len(np.atleast_1d(pub[2])):
                raise ValueError("parameterised circuits are not supported: bind the parameters first")
            prec = float(pub[3]) if len(pub) > 3 and pub[3] is not None else (precision or self.default_precision)
            results.append(self._run_pub(pub[0], pub[1], prec))
        return EstimatorJob(results)

    def _run_pub(self, circuit, observable, precision):
        circ = as_circuit(circuit)
        terms = observable_terms(observable)
        if not terms:
            raise ValueError("empty observable")
        width = len(terms[0][0])
        if any(len(lbl) != width for lbl, _ in terms):
            raise ValueError("all Pauli labels of an observable must have the same length")
        if width > circ.num_qubits:
            raise ValueError(f"observable acts on {width} qubits, the circuit has {circ.num_qubits}")
        shots = int(math.ceil(1.0 / precision ** 2))
        const = sum(c for lbl, c in terms if set(lbl) == {"I"})
        active = [(lbl + "I" * 0, c) for lbl, c in terms if set(lbl) != {"I"}]
        groups = group_qubitwise_commuting(active) if self.abelian_grouping else \
            [(lbl[::-1], [k]) for k, (lbl, _c) in enumerate(active)]
        circuits, layouts = [], []
        for basis, _members in groups:
            mc, qubits = measurement_circuit(circ, basis + "I" * (circ.num_qubits - width))
            circuits.append(mc)
            layouts.append({q: i for i, q in enumerate(qubits)})
        ev, var = const, 0.0
        if circuits:
            kw = {"shots": shots}
            if self.seed_simulator is not None:
                kw["seed_simulator"] = self.seed_simulator
            res = self.backend.run(circuits, **kw).result()
            for gi, (basis, members) in enumerate(groups):
                counts = res.get_counts(gi)
                for k in members:
                    lbl, c = active[k]
                    pos = [layouts[gi][q] for q, ch in enumerate(lbl[::-1]) if ch != "I"]
                    e = pauli_expectation(counts, pos)
                    ev += c * e
                    var += c * c * max(0.0, 1.0 - e * e) / shots
        data = _DataBin(np.float64(ev), np.float64(math.sqrt(var)), shots)
        return PubResult(data, {"target_precision": precision, "shots": shots, "circuits": len(circuits),
                                "groups": [b for b, _ in groups]})


# ----------------------------------------------------------------------------------- sampler primitive (dtc_qasm.py:138-140)
class BitArray:
    """What `result[0].data.<register>` offers in the scripts: get_counts(), num_shots, num_bits."""

    def __init__(self, counts, num_bits):
        self._counts, self.num_bits = dict(counts), int(num_bits)
        self.num_shots = int(sum(counts.values()))

    def get_counts(self):
        return dict(self._counts)

    def get_int_counts(self):
        return {int(k.replace(" ", ""), 2): v for k, v in self._counts.items()}


class _SamplerData:
    """`data.c`, `data.meas`, ...: this container's circuits carry one classical register, whatever its name."""

    def __init__(self, bits):
        self._bits = bits

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return self._bits

    def values(self):
        return [self._bits]


class BackendSamplerV2:
    """Same call surface as qiskit.primitives.BackendSamplerV2 / qiskit_ibm_runtime.SamplerV2 for what dtc_qasm.py uses:
        sampler.run([circuit], shots=1024).result()[0].data.c.get_counts()          (dtc_qasm.py:138-140)
    All pubs of a call go through one pipelined `backend.run(list, shots)`.  A pub is a circuit (native, OpenQASM-2 text or
    qiskit-like) or a tuple (circuit, parameter_values=None, shots=None); per-pub shot counts split the call."""

    def __init__(self, backend=None, mode=None, options=None):
        self.backend = backend if backend is not None else mode
        if self.backend is None:
            raise ValueError("BackendSamplerV2 needs a backend")
        opts = dict(options or {})
        self.default_shots = int(opts.get("default_shots", 1024))
        self.seed_simulator = opts.get("seed_simulator")

    def run(self, pubs, shots=None):
        circs, nshots = [], []
        for pub in pubs:
            if isinstance(pub, (tuple, list)):
                if len(pub) > 1 and pub[1] is not None and len(np.atleast_1d(pub[1])):
                    raise ValueError("parameterised circuits are not supported: bind the parameters first")
                circs.append(as_circuit(pub[0]))
                nshots.append(int(pub[2]) if len(pub) > 2 and pub[2] is not None else None)
            else:
                circs.append(as_circuit(pub))
                nshots.append(None)
        nshots = [n if n is not None else (int(shots) if shots is not None else self.default_shots) for n in nshots]
        results = [None] * len(circs)
        for n in sorted(set(nshots)):
            idx = [i for i, m in enumerate(nshots) if m == n]
            kw = {"shots": n}
            if self.seed_simulator is not None:
                kw["seed_simulator"] = [int(self.seed_simulator) + i for i in idx]
            res = self.backend.run([circs[i] for i in idx], **kw).result()
            for j, i in enumerate(idx):
                results[i] = PubResult(_SamplerData(BitArray(res.get_counts(j), circs[i].num_clbits)), {"shots": n})
        return EstimatorJob(results)


SamplerV2 = BackendSamplerV2
