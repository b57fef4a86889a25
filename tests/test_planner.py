"""Host logic: ingestion, lowering, noise adaptor, layer compiler (checked with the numpy interpreter)."""
import math

import numpy as np
import pytest

import dtcsim
import program_interp as PI
import refcircuits as RC
from dtcsim import QuantumCircuit, compile_circuit
from oracle import dtc_circuits as C
from oracle import oracle as O


def _oracle_ops(circ):
    return O.compact_ops(RC.ops_of(circ), circ.num_qubits)


@pytest.mark.parametrize("pol", ["x", "y", "xy", "yx"])
@pytest.mark.parametrize("state", ["vacuum", "neel"])
def test_lowering_matches_oracle_restatement(disorder, pol, state):
    """dtcsim's builder + pass-manager mirror emits the op stream the oracle restates from fast.py."""
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    for t, echo in ((0, False), (2, False), (2, True)):
        circ = RC.transpiled(RC.qc_body(state, 4, 0.84, hs, phis, t, 2, echo, pol))
        ops, _, _ = C.autocorr_gates(state, 4, 0.84, hs, phis, t, 2, echo, pol)
        want = C.lower_level0(ops, C.SNAKE_LAYOUT)
        got = RC.ops_of(circ)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert a[0] == b[0] and tuple(a[1]) == tuple(b[1]) and tuple(a[3]) == tuple(b[3])
            assert np.allclose(a[2], b[2], atol=0, rtol=0)
        assert circ.num_qubits == 31


def test_gate_counts_via_dtcsim_api(disorder, gate_counts):
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    for key in ("t0_forward", "t1_forward", "t1_echo", "t20_echo"):
        t, kind = key.split("_")
        circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, int(t[1:]), 2, kind == "echo"))
        assert circ.count_ops() == gate_counts["L4"][key]


def test_time_dependent_g_circuits(disorder):
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    gv = [0.84, 0.85, 0.9]
    circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, 3, 2, True, g_values=gv))
    ops, _, _ = C.autocorr_gates("vacuum", 4, 0.84, hs, phis, 3, 2, True, g_values=gv)
    want = C.lower_level0(ops, C.SNAKE_LAYOUT)
    assert [o[0] for o in RC.ops_of(circ)] == [o[0] for o in want]
    prog = compile_circuit(circ, None)
    oc, na, _ = _oracle_ops(circ)
    assert np.abs(PI.to_circuit_order(PI.run(prog), prog)[0] - O.run_statevector(oc, na)).max() < 1e-13


@pytest.mark.parametrize("pol,t,echo", [("x", 3, False), ("x", 2, True), ("xy", 2, True), ("y", 2, False)])
def test_program_semantics_reference_circuits(disorder, pol, t, echo):
    """Planner output executed by the numpy interpreter == gate-by-gate oracle (ideal, trajectories, DM)."""
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    circ = RC.transpiled(RC.qc_body("neel", 4, 0.84, hs, phis, t, 2, echo, pol))
    nm, onoise = RC.noise_model(0.05), O.PauliNoise.depolarizing(0.05)
    oc, na, act = _oracle_ops(circ)
    prog0 = compile_circuit(circ, None)
    assert prog0.n == na == 5 and prog0.active == act
    assert np.abs(PI.to_circuit_order(PI.run(prog0), prog0)[0] - O.run_statevector(oc, na)).max() < 1e-13
    prog = compile_circuit(circ, nm, want_dm=True)
    assert prog.n_sites == len(O.noise_sites(oc, onoise))
    trajs = np.arange(64)
    assert np.abs(PI.to_circuit_order(PI.run(prog, 4321, trajs), prog) - O.run_trajectories(oc, na, onoise, 4321, trajs)).max() < 1e-12
    assert np.abs(PI.dm_to_circuit_order(PI.run_dm(prog, prog.n), prog) - O.run_density_matrix(oc, na, onoise)).max() < 1e-12


def _random_circuit(rng, n, depth, measure=True):
    c = QuantumCircuit(n, n)
    names1 = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "sxdg", "id"]
    for _ in range(depth):
        r = rng.integers(0, 10)
        q = int(rng.integers(0, n))
        if r == 0:
            getattr(c, names1[int(rng.integers(0, len(names1)))])(q)
        elif r == 1:
            c.u3(*rng.uniform(-4, 4, 3), q)
        elif r == 2:
            c.u2(*rng.uniform(-4, 4, 2), q)
        elif r == 3:
            c.u1(rng.uniform(-4, 4), q)
        elif r == 4:
            getattr(c, ["rx", "ry", "rz"][int(rng.integers(0, 3))])(rng.uniform(-7, 7), q)
        else:
            q2 = int(rng.integers(0, n - 1))
            q2 = q2 + 1 if q2 >= q else q2
            kind = int(rng.integers(0, 4))
            if kind == 0:
                c.cx(q, q2)
            elif kind == 1:
                c.cz(q, q2)
            elif kind == 2:
                c.rzz(rng.uniform(-4, 4), q, q2)
            else:
                c.swap(q, q2)
    if measure:
        c.measure_all()
    return c


@pytest.mark.parametrize("seed", range(6))
def test_program_semantics_random_circuits(seed):
    """Arbitrary gates from the supported set, noise on u1/u2/u3/h/rx: ideal + trajectories + DM."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 6))
    circ = _random_circuit(rng, n, 60)
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.2, 1), ["u1", "u2", "u3"])
    nm.add_all_qubit_quantum_error(dtcsim.pauli_error([("X", 0.1), ("Z", 0.15), ("I", 0.75)]), ["h", "rx"])
    onoise = O.PauliNoise({"u1": (0.05, 0.05, 0.05), "u2": (0.05, 0.05, 0.05), "u3": (0.05, 0.05, 0.05),
                           "h": (0.1, 0.0, 0.15), "rx": (0.1, 0.0, 0.15)})
    oc, na, _ = _oracle_ops(circ)
    prog0 = compile_circuit(circ, None)
    assert np.abs(PI.to_circuit_order(PI.run(prog0), prog0)[0] - O.run_statevector(oc, na)).max() < 1e-12
    prog = compile_circuit(circ, nm, want_dm=True)
    trajs = np.arange(32)
    assert np.abs(PI.to_circuit_order(PI.run(prog, seed, trajs), prog) - O.run_trajectories(oc, na, onoise, seed, trajs)).max() < 1e-12
    assert np.abs(PI.dm_to_circuit_order(PI.run_dm(prog, prog.n), prog) - O.run_density_matrix(oc, na, onoise)).max() < 1e-12


def test_cx_rz_cx_peephole_restores_rzz(disorder):
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", 20, 0.97, hs, phis, 4, 10, False))
    prog = compile_circuit(circ, RC.noise_model())
    # one rotation layer per period plus the skewed ancilla prelude/postlude
    assert prog.n == 21 and prog.n_layers <= 4 + 8
    n_rot = sum(1 for t in prog.ev_type if t == 0)
    assert n_rot == 4 * 20 + 10          # 20 rx per period + 10 ancilla rotations (6 u2 + 2 x (cx -> 2 ry))


def test_idle_qubit_truncation_and_clbits():
    c = QuantumCircuit(40, 3)
    c.h(33)
    c.cx(33, 7)
    c.measure(7, 2)
    c.measure(33, 0)
    prog = compile_circuit(c, None)
    assert prog.n == 2 and prog.active == [7, 33]
    assert prog.measures == [(prog.bit_of[1], 0), (prog.bit_of[0], 2)]
    assert sorted(prog.order) == [0, 1]


def test_error_behaviour():
    c = QuantumCircuit(2, 1)
    c.h(0)
    c.measure(0, 0)
    c.x(0)
    with pytest.raises(ValueError):
        compile_circuit(c, None)                       # gate after measurement
    c2 = QuantumCircuit(2, 1)
    c2._add("ccx_like", (0, 1))
    with pytest.raises(ValueError):
        compile_circuit(c2, None)                      # unsupported instruction
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.1, 1), ["cx"])
    c3 = QuantumCircuit(2, 1)
    c3.cx(0, 1)
    c3.measure(1, 0)
    with pytest.raises(ValueError):
        compile_circuit(c3, nm)                        # noise on a 2q gate
    with pytest.raises(ValueError):
        dtcsim.depolarizing_error(0.1, 2)
    with pytest.raises(ValueError):
        QuantumCircuit(2).cx(0, 0)
    with pytest.raises(ValueError):
        QuantumCircuit(2).h(5)
    with pytest.raises(ValueError):
        compile_circuit(QuantumCircuit(3, 1), None)    # empty circuit


def test_noise_model_from_qiskit_dict():
    """Shape of qiskit_aer NoiseModel.to_dict() for depolarizing_error(0.05,1) on u1/u2/u3."""
    p = 0.05
    d = {"errors": [{"type": "qerror", "operations": ["u1", "u2", "u3"],
                     "instructions": [[{"name": "id", "qubits": [0]}], [{"name": "x", "qubits": [0]}],
                                      [{"name": "y", "qubits": [0]}], [{"name": "z", "qubits": [0]}]],
                     "probabilities": [1 - 3 * p / 4, p / 4, p / 4, p / 4]}]}

    class Fake:
        def to_dict(self):
            return d

    nm = dtcsim.as_noise_model(Fake())
    assert nm.lookup("u3", 5) == (p / 4, p / 4, p / 4) and nm.lookup("rz", 0) is None
    d2 = {"errors": [{"type": "qerror", "operations": ["u3"], "gate_qubits": [[2]],
                      "instructions": [[{"name": "pauli", "params": ["X"], "qubits": [0]}], [{"name": "id", "qubits": [0]}]],
                      "probabilities": [0.3, 0.7]}]}
    nm2 = dtcsim.as_noise_model(d2)
    assert nm2.lookup("u3", 2) == (0.3, 0.0, 0.0) and nm2.lookup("u3", 1) is None
    # an identity readout error is an ideal model; thermal-relaxation-like (non-Pauli) entries still raise
    assert dtcsim.as_noise_model({"errors": [{"type": "roerror", "operations": ["measure"], "probabilities": [[1, 0], [0, 1]]}]}) is None
    with pytest.raises(ValueError):
        dtcsim.as_noise_model({"errors": [{"type": "qerror", "operations": ["u3"], "probabilities": [1.0],
                                           "instructions": [[{"name": "kraus", "qubits": [0], "params": []}]]}]})
    assert dtcsim.as_noise_model(dtcsim.NoiseModel()) is None
    # composing errors on the same instruction (energy.py:214-218 loop)
    nm3 = dtcsim.NoiseModel()
    nm3.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.1, 1), ["u3"])
    nm3.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.2, 1), ["u3"])
    px, py, pz = nm3.lookup("u3", 0)
    lam = (1 - 0.1) * (1 - 0.2)                          # depolarizing channels compose multiplicatively
    assert abs(px - (1 - lam) / 4) < 1e-15 and px == py == pz


def test_qasm2_ingest_dtc_dialect(disorder):
    """dtc_qasm.py:95-107 round trip: the OpenQASM-2 text PennyLane writes for the L-qubit circuit."""
    hs, phis = disorder[20][0][0][:6], disorder[20][1][0][:5]
    ops, n, nc = C.dtc_qasm_gates("1", 6, 0.94, hs, phis, 2)
    lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', "qreg q[6];", "creg c[6];"]
    for name, qs, params, cs in ops:
        if name == "measure":
            lines.append(f"measure q[{qs[0]}] -> c[{cs[0]}];")
        else:
            arg = ",".join(f"q[{q}]" for q in qs)
            par = f"({','.join(repr(p) for p in params)})" if params else ""
            lines.append(f"{name}{par} {arg};")
    circ = dtcsim.from_qasm2("\n".join(lines))
    assert [o.astuple() for o in circ.ops] == [(a, tuple(b), tuple(c), tuple(d)) for a, b, c, d in ops]
    prog = compile_circuit(circ, None)
    assert np.abs(PI.to_circuit_order(PI.run(prog), prog)[0] - O.run_statevector(ops, 6)).max() < 1e-13
    c2 = dtcsim.from_qasm2('OPENQASM 2.0; qreg q[2]; creg c[2]; rx(pi/2) q[0]; h q; cx q[0],q[1]; measure q -> c;')
    assert [o.name for o in c2.ops] == ["rx", "h", "h", "cx", "measure", "measure"]
    assert abs(c2.ops[0].params[0] - math.pi / 2) < 1e-15


def test_from_qiskit_duck_typing():
    class Bit:
        pass

    class Opn:
        def __init__(self, name, params):
            self.name, self.params = name, params

    class Inst:
        def __init__(self, op, qs, cs):
            self.operation, self.qubits, self.clbits = op, qs, cs

    class Loc:
        def __init__(self, i):
            self.index = i

    class FakeQC:
        def __init__(self):
            self.qb = [Bit() for _ in range(3)]
            self.cb = [Bit()]
            self.num_qubits, self.num_clbits, self.name, self.global_phase = 3, 1, "fake", 0.0
            self.data = [Inst(Opn("u3", [0.1, 0.2, 0.3]), [self.qb[2]], []), Inst(Opn("barrier", []), self.qb, []),
                         Inst(Opn("cx", []), [self.qb[2], self.qb[0]], []), Inst(Opn("measure", []), [self.qb[0]], [self.cb[0]])]

        def find_bit(self, b):
            return Loc((self.qb + self.cb).index(b) if b in self.qb else self.cb.index(b))

    c = dtcsim.as_circuit(FakeQC())
    assert [o.astuple() for o in c.ops] == [("u3", (2,), (0.1, 0.2, 0.3), ()), ("cx", (2, 0), (), ()),
                                            ("measure", (0,), (), (0,))]


def test_inverse_is_inverse():
    rng = np.random.default_rng(3)
    c = _random_circuit(rng, 3, 40, measure=False)
    full = QuantumCircuit(3)
    full.append(c, range(3))
    full.append(c.inverse(), range(3))
    psi = O.run_statevector(RC.ops_of(full), 3)
    assert abs(abs(psi[0]) - 1) < 1e-12


def test_backend_program_cache(disorder):
    """run() compiles a circuit once per (op list, noise model): a second run of the same circuit reuses the program."""
    import refcircuits as RC
    hs, phis = disorder[20][0][0][:6], disorder[20][1][0][:5]
    c1 = RC.transpiled(RC.qc_body("vacuum", 6, 0.97, hs, phis, 2, 3, True))
    c2 = RC.transpiled(RC.qc_body("vacuum", 6, 0.97, hs, phis, 2, 3, True))       # equal content, different object
    c3 = RC.transpiled(RC.qc_body("vacuum", 6, 0.84, hs, phis, 2, 3, True))
    nm = dtcsim.as_noise_model(RC.noise_model(0.05))
    nm2 = dtcsim.as_noise_model(RC.noise_model(0.1))
    sim = dtcsim.AerSimulator()
    p1 = sim._compiled(dtcsim.as_circuit(c1), nm)
    assert sim._compiled(dtcsim.as_circuit(c2), nm) is p1
    assert sim._compiled(dtcsim.as_circuit(c3), nm) is not p1
    assert sim._compiled(dtcsim.as_circuit(c1), nm2) is not p1
    assert sim._compiled(dtcsim.as_circuit(c1), None) is not p1
    assert len(sim._prog_cache) == 4


def test_readout_error_model_and_host_philox():
    """Readout errors (fast.py:77-78 device-calibrated noise, classical part): parsing, lookup, host Philox == oracle Philox."""
    import dtcsim
    from dtcsim import noise as N
    from dtcsim import philox_np
    from oracle import philox
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_readout_error(dtcsim.ReadoutError([[0.97, 0.03], [0.08, 0.92]]))
    nm.add_readout_error([[0.9, 0.1], [0.2, 0.8]], [5])
    assert nm.has_readout_noise() and not nm.has_gate_noise() and not nm.is_ideal()
    assert nm.lookup_readout(5) == [[0.9, 0.1], [0.2, 0.8]] and nm.lookup_readout(0)[1][0] == 0.08
    d = {"errors": [{"type": "roerror", "operations": ["measure"], "probabilities": [[0.99, 0.01], [0.05, 0.95]],
                     "gate_qubits": [[3]]},
                    {"type": "qerror", "operations": ["u3"], "instructions": [[{"name": "x", "qubits": [0]}], [{"name": "id", "qubits": [0]}]],
                     "probabilities": [0.1, 0.9]}]}
    nm2 = N.as_noise_model(d)
    assert nm2.lookup_readout(3) == [[0.99, 0.01], [0.05, 0.95]] and nm2.lookup_readout(2) is None and nm2.has_gate_noise()
    with pytest.raises(ValueError):
        dtcsim.ReadoutError([[0.5, 0.6], [0.1, 0.9]])
    idx = np.array([0, 1, 7, 2 ** 31 + 5])
    traj = np.array([0, 3, 2 ** 33 + 1, 9])
    for seed in (0, 1234, 2 ** 40 + 3):
        assert np.array_equal(philox_np.uniform(seed, idx, 2, traj), philox.uniform(seed, idx, 2, traj))
    # flips: deterministic under the contract, and the recorded distribution is the assignment matrix applied to the true one
    from dtcsim.backend import DTCSimulator
    vals = np.zeros(200000, dtype=np.int64)
    rec = DTCSimulator._readout_flips(vals, {0: [[0.97, 0.03], [0.08, 0.92]], 2: [[0.5, 0.5], [0.0, 1.0]]}, 9, np.arange(200000))
    assert abs(((rec >> 0) & 1).mean() - 0.03) < 0.002 and abs(((rec >> 2) & 1).mean() - 0.5) < 0.005 and ((rec >> 1) & 1).sum() == 0
    pr = DTCSimulator._readout_probs({0: 0.25, 1: 0.75}, {0: [[0.97, 0.03], [0.08, 0.92]]})
    assert abs(pr[0] - (0.25 * 0.97 + 0.75 * 0.08)) < 1e-15 and abs(pr[0] + pr[1] - 1) < 1e-15
