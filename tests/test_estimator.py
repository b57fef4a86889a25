"""Estimator path (energy.py:83-102,166-171): Pauli grouping, measurement circuits, label order, evs from counts.
CPU part: a fake backend whose counts come from the exact oracle distribution; GPU part: the real backend vs the oracle."""
import numpy as np
import pytest

import dtcsim
from dtcsim import estimator as E
from oracle import dtc_circuits as C
from oracle import oracle as O


def _energy_circuit(L, g, hs, phis, t):
    """energy.py:123-135: L qubits, no ancilla, t periods."""
    c = dtcsim.QuantumCircuit(L)
    for _ in range(t):
        for i in range(L):
            c.rx(np.pi * g, i)
        for i in range(0, L - 1, 2):
            c.rzz(phis[i], i, i + 1)
        for i in range(1, L - 1, 2):
            c.rzz(phis[i], i, i + 1)
        for i in range(L):
            c.rz(hs[i], i)
    return c


def _exact_energy(ops, L, terms, noise=None):
    """<H> from the oracle's density matrix; labels in qiskit order (rightmost character = qubit 0)."""
    low = C.lower_level0(ops)
    rho = O.run_density_matrix(low, L, noise)
    paulis = {"I": np.eye(2), "X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1.0, -1.0])}
    tot = 0.0
    for label, c in terms:
        op = np.array([[1.0]])
        for ch in label:                      # leftmost character = highest qubit = most significant kron factor
            op = np.kron(op, paulis[ch])
        tot += c * np.real(np.trace(rho @ op))
    return float(tot)


def test_hamiltonian_labels_match_reference_quirk():
    L = 4
    hs, phis = [0.1, 0.2, 0.3, 0.4], [1.0, 2.0, 3.0]
    ham = E.dtc_hamiltonian(L, 0.9, phis, hs)
    assert ham[0] == ("ZIII", 0.1) and ham[3] == ("IIIZ", 0.4)            # hs[0] sits on qubit L-1 (energy.py:89-91)
    assert ham[4] == ("ZZII", 1.0) and ham[6] == ("IIZZ", 3.0)
    assert ham[7][0] == "XIII" and abs(ham[7][1] - 0.9 * np.pi) < 1e-15
    assert len(ham) == 3 * L - 1


def test_grouping_and_measurement_circuits():
    ham = E.dtc_hamiltonian(5, 0.97, [1, 2, 3, 4], [1, 2, 3, 4, 5])
    groups = E.group_qubitwise_commuting(ham)
    assert [b for b, _ in groups] == ["ZZZZZ", "XXXXX"]                   # all Z / ZZ terms share one circuit, the X terms another
    assert sorted(groups[0][1]) == list(range(9)) and sorted(groups[1][1]) == list(range(9, 14))
    c = dtcsim.QuantumCircuit(5)
    c.rx(0.3, 2)
    mc, qubits = E.measurement_circuit(c, "XIYZI")
    names = [(o.name, o.qubits) for o in mc.ops]
    assert qubits == [0, 2, 3] and names == [("rx", (2,)), ("h", (0,)), ("sdg", (2,)), ("h", (2,)), ("measure", (0,)),
                                             ("measure", (2,)), ("measure", (3,))]
    assert E.pauli_expectation({"011": 3, "100": 1}, [0, 1]) == 1.0 and E.pauli_expectation({"01": 1, "00": 3}, [0]) == 0.5
    with pytest.raises(ValueError):
        c.measure_all() if hasattr(c, "measure_all") else None
        c2 = dtcsim.QuantumCircuit(1, 1)
        c2.measure(0, 0)
        E.measurement_circuit(c2, "Z")


class _OracleBackend:
    """Counts drawn from the oracle's exact outcome distribution (stands in for run() on the CPU)."""

    def __init__(self, noise=None, seed=5):
        self.noise, self.rng = noise, np.random.default_rng(seed)

    def run(self, circuits, shots=1024, **_kw):
        outs = []
        for circ in circuits:
            ops = [o.astuple() for o in dtcsim.lower_level0(circ).ops]
            n, ncl = circ.num_qubits, circ.num_clbits
            rho = O.run_density_matrix(ops, n, self.noise)
            pr = O.outcome_probabilities(np.real(np.diag(rho)).copy(), n, O.measured_map(ops), ncl)
            vals = self.rng.choice(len(pr), size=shots, p=pr / pr.sum())
            outs.append(O.counts_dict(vals, ncl))
        backend = self

        class _R:
            def get_counts(self, i):
                return outs[i]

        class _J:
            def result(self):
                return _R()
        return _J()


@pytest.mark.parametrize("noisy", [False, True])
def test_estimator_energy_vs_exact(disorder, noisy):
    L, t, g = 5, 3, 0.97
    hs, phis = disorder[20][0][0][:L], disorder[20][1][0][:L - 1]
    circ = _energy_circuit(L, g, hs, phis, t)
    ham = E.dtc_hamiltonian(L, g, phis, hs)
    noise = O.PauliNoise.depolarizing(0.05) if noisy else None
    est = E.BackendEstimatorV2(_OracleBackend(noise))
    pub = est.run([(circ, ham)]).result()[0]
    exact = _exact_energy([o.astuple() for o in circ.ops], L, ham, noise)
    assert pub.metadata["shots"] == 4096 and pub.metadata["circuits"] == 2
    assert abs(float(pub.data.evs) - exact) < 4.5 * float(pub.data.stds) + 1e-12
    # per-pub precision -> shot count, identity terms are constants
    pub2 = est.run([(circ, [("IIIII", 2.5), ("IIIIZ", 1.0)], None, 1 / 32)]).result()[0]
    assert pub2.metadata["shots"] == 1024 and abs(float(pub2.data.evs) - 2.5) <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("L,noisy", [(5, True), (6, False), (14, True)])
def test_estimator_on_gpu_vs_exact(disorder, L, noisy):
    """energy.py:166-171 through the real backend: <H> within the estimator's own standard error of the exact value (L = 14:
    the measured register is wider than the 12-qubit probability path, i.e. per-trajectory basis-state samples)."""
    from conftest import family_z
    t, g = 2, 0.97
    hs, phis = disorder[20][0][0][:L], disorder[20][1][0][:L - 1]
    circ = _energy_circuit(L, g, hs, phis, t)
    ham = E.dtc_hamiltonian(L, g, phis, hs)
    nm = None
    if noisy:
        nm = dtcsim.NoiseModel()
        nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u1", "u2", "u3"], warnings=False)
    sim = dtcsim.AerSimulator(noise_model=nm, device="GPU", cuStateVec_enable=True)
    pm = dtcsim.generate_preset_pass_manager(optimization_level=0, backend=sim)
    est = dtcsim.BackendEstimatorV2(backend=sim, options={"seed_simulator": 7})
    pub = est.run([(pm.run(circ), ham)]).result()[0]
    if L <= 6:
        exact = _exact_energy([o.astuple() for o in circ.ops], L, ham, O.PauliNoise.depolarizing(0.05) if noisy else None)
    else:                                                   # trajectory average of the oracle's statevectors
        low = C.lower_level0([o.astuple() for o in circ.ops])
        psi = O.run_trajectories(low, L, O.PauliNoise.depolarizing(0.05), 11, np.arange(600))
        idx = np.arange(1 << L)
        exact, p = 0.0, (np.abs(psi) ** 2).mean(0)
        for label, c in ham:
            word = label[::-1]
            if "X" in word:
                q = word.index("X")
                exact += c * float(np.mean(np.real(np.sum(np.conj(psi) * psi[:, idx ^ (1 << q)], axis=1))))
            else:
                z = np.ones(1 << L)
                for q, ch in enumerate(word):
                    if ch == "Z":
                        z = z * (1 - 2 * ((idx >> q) & 1))
                exact += c * float(np.dot(p, z))
    tol = family_z(3) * float(pub.data.stds) + (0.0 if L <= 6 else 0.15)   # + Monte-Carlo error of the 600-trajectory reference
    assert abs(float(pub.data.evs) - exact) < tol, (float(pub.data.evs), exact, float(pub.data.stds))


def test_estimator_y_and_mixed_terms(disorder):
    """Basis changes for Y (sdg, h) and qubit-wise commuting groups of mixed X / Y / Z words, against exact values."""
    L, g = 4, 0.6
    hs, phis = disorder[20][0][1][:L], disorder[20][1][1][:L - 1]
    circ = _energy_circuit(L, g, hs, phis, 2)
    terms = [("IIYI", 0.7), ("XYIZ", -1.3), ("ZZII", 0.4), ("IYXI", 0.9), ("YIIY", 0.5), ("IIII", 0.25)]
    groups = E.group_qubitwise_commuting([t for t in terms if set(t[0]) != {"I"}])
    assert len(groups) <= 4
    est = E.BackendEstimatorV2(_OracleBackend(None, seed=11), options={"default_precision": 1 / 128})
    pub = est.run([(circ, terms)]).result()[0]
    exact = _exact_energy([o.astuple() for o in circ.ops], L, terms)
    assert pub.metadata["shots"] == 16384
    assert abs(float(pub.data.evs) - exact) < 4.5 * float(pub.data.stds) + 1e-12, (float(pub.data.evs), exact)


def test_run_energy_sweep_vs_exact_energy(disorder):
    """sweeps.run_energy_sweep (energy.py:173-195 loops: instance x t, <H>/L per point through the estimator) on the
    oracle-backed stand-in simulator: every point within 5 standard errors of Tr(rho H)/L of the noisy density matrix, the
    t = 0 Z-part exact, shape / bookkeeping as documented."""
    from test_dist_cpu import _OracleSim
    L, g = 4, 0.97
    hs, phis = disorder[20][0][:2, :L], disorder[20][1][:2, :L - 1]
    tv = [0, 1, 3]
    res = dtcsim.run_energy_sweep(_OracleSim(), L, g, hs, phis, tv, precision=1 / 32, seed_simulator=3)
    assert res["energy_per_site"].shape == (2, 3) and res["points"] == 6 and res["circuits"] == 12
    assert np.allclose(res["mean"], res["energy_per_site"].mean(axis=0))
    noise = O.PauliNoise.depolarizing(0.05)
    for i in range(2):
        ham = E.dtc_hamiltonian(L, g, phis[i], hs[i])
        for k, t in enumerate(tv):
            ops = [o.astuple() for o in dtcsim.energy_circuit(L, g, hs[i], phis[i], t, transpile=False).ops]
            want = _exact_energy(ops, L, ham, noise) / L
            assert abs(res["energy_per_site"][i, k] - want) < 5 * max(res["stds"][i, k], 1e-3), (i, t)
    # echo circuits: t periods forward and t back
    c = dtcsim.energy_circuit(L, g, hs[0], phis[0], 2, echo=True, transpile=False)
    assert len(c.ops) == 4 * (L + (L - 1) + L)


def test_backend_sampler_v2_dtc_qasm_usage(disorder):
    """dtc_qasm.py:138-140: `SamplerV2(mode=backend).run([qc], shots=1024).result()[0].data.c.get_counts()` with the circuit
    handed over as OpenQASM-2 text; counts are the backend's (here the oracle-backed stand-in), per-pub shot counts honoured."""
    from test_dist_cpu import _OracleSim

    class _Sim(_OracleSim):
        def run(self, circuits, **kw):
            from dtcsim.ir import as_circuit
            return super().run([as_circuit(c) for c in circuits], **kw)

    hs, phis = disorder[20][0][0][:4], disorder[20][1][0][:3]
    c1 = dtcsim.expz_circuit(4, 0.94, hs, phis, 2, "1")
    c2 = dtcsim.expz_circuit(4, 0.94, hs, phis, 3, "0")
    sampler = dtcsim.SamplerV2(mode=_Sim(), options={"seed_simulator": 12})
    res = sampler.run([c1.qasm(), (c2, None, 64)], shots=256).result()
    a, b = res[0].data.c, res[1].data.meas
    assert a.num_shots == 256 and b.num_shots == 64 and a.num_bits == 4
    noise = O.PauliNoise.depolarizing(0.05)
    want = O.run_counts([o.astuple() for o in c1.ops], 4, 4, shots=256, noise=noise, seed=12)[0]
    assert a.get_counts() == want and sum(a.get_int_counts().values()) == 256
    want2 = O.run_counts([o.astuple() for o in c2.ops], 4, 4, shots=64, noise=noise, seed=13)[0]
    assert b.get_counts() == want2
    with pytest.raises(ValueError):
        dtcsim.BackendSamplerV2()
