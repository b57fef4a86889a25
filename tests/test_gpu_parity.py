"""GPU parity tests: the CUDA library (through its C ABI / the AerSimulator-compatible backend) against
the CPU oracle on identical inputs.  Tolerances (BASELINE.json north_star): noiseless and per-trajectory
statevector amplitudes 1e-10, density-matrix expectation values 1e-8, shot/trajectory estimates within 3 sigma --
held family-wise over the points of a test (conftest.family_z: per-point bound 3.0 for one point, 3.5 for six).
"""
import numpy as np
import pytest

import dtcsim
import program_interp as PI
import refcircuits as RC
from conftest import family_z, golden_csv
from dtcsim import compile_circuit
from oracle import c_oracle as CO
from oracle import dtc_circuits as C
from oracle import oracle as O
from test_planner import _random_circuit

pytestmark = pytest.mark.gpu

AMP_TOL = 1e-10
DM_TOL = 1e-8


@pytest.fixture(scope="module")
def ctx():
    from dtcsim import backend
    return backend.DeviceContext(0)


def _evolve_true(ctx, prog, n_traj, offset, seed, engine=0):
    from dtcsim import backend
    batch = backend.evolve(ctx, prog, n_traj, offset, seed, engine=engine)
    psi = batch.materialize().cpu().numpy()
    return PI.to_circuit_order(psi, prog), batch


def test_frames_bit_exact_vs_spec(ctx, disorder):
    """Philox4x32-10 Pauli sampling + frame walk on the device == the numpy specification, bit for bit."""
    from dtcsim import backend
    hs, phis = disorder[20][0][0][:12], disorder[20][1][0][:11]
    circ = RC.transpiled(RC.qc_body("neel", 12, 0.97, hs, phis, 6, 6, True, "xy"))
    prog = compile_circuit(circ, RC.noise_model(0.2))
    batch = backend.evolve(ctx, prog, 300, 1000, 424242)
    fx, fz, ph = batch.frames_host()
    layers, k_of = PI.build_layers(prog)
    _, rfx, rfz, rph = PI.frame_walk(prog, k_of, 424242, np.arange(1000, 1300))
    assert np.array_equal(fx, rfx) and np.array_equal(fz, rfz) and np.array_equal(ph, rph.astype(np.int32))
    assert (fx != 0).any() and (fz != 0).any()


@pytest.mark.parametrize("L,t,echo,pol,layout,state", [
    (11, 2, False, "x", True, "vacuum"),
    (12, 3, True, "x", False, "neel"),
    (13, 2, True, "xy", True, "vacuum"),
    (15, 2, False, "y", True, "neel"),
])
def test_tile_engine_trajectories_vs_oracle(ctx, disorder, L, t, echo, pol, layout, state):
    hs, phis = disorder[20][0][1][:L], disorder[20][1][1][:L - 1]
    circ = RC.transpiled(RC.qc_body(state, L, 0.97, hs, phis, t, L // 2, echo, pol), layout=layout)
    oc, na, _ = O.compact_ops(RC.ops_of(circ), circ.num_qubits)
    prog = compile_circuit(circ, RC.noise_model(0.05))
    psi, _ = _evolve_true(ctx, prog, 6, 40, 77, engine=2)
    ref = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(0.05), 77, np.arange(40, 46))
    assert np.abs(psi - ref).max() < AMP_TOL
    prog0 = compile_circuit(circ, None)
    psi0, _ = _evolve_true(ctx, prog0, 1, 0, 0, engine=2)
    assert np.abs(psi0[0] - O.run_statevector(oc, na)).max() < AMP_TOL


def test_generic_engine_vs_oracle(ctx, disorder):
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, 5, 2, True))
    oc, na, _ = O.compact_ops(RC.ops_of(circ), 31)
    prog = compile_circuit(circ, RC.noise_model(0.05))
    psi, _ = _evolve_true(ctx, prog, 64, 0, 9)
    ref = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(0.05), 9, np.arange(64))
    assert np.abs(psi - ref).max() < AMP_TOL


@pytest.mark.parametrize("seed", range(4))
def test_random_circuits(ctx, seed):
    rng = np.random.default_rng(200 + seed)
    n = 12 + seed
    circ = _random_circuit(rng, n, 80)
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.3, 1), ["u1", "u2", "u3", "h"])
    onoise = O.PauliNoise.depolarizing(0.3, names=("u1", "u2", "u3", "h"))
    oc, na, _ = O.compact_ops(RC.ops_of(circ), n)
    prog = compile_circuit(circ, nm)
    psi, _ = _evolve_true(ctx, prog, 3, 5, seed, engine=2)
    ref = O.run_trajectories(oc, na, onoise, seed, np.arange(5, 8))
    assert np.abs(psi - ref).max() < AMP_TOL
    psi_g, _ = _evolve_true(ctx, prog, 3, 5, seed, engine=1)
    assert np.abs(psi_g - ref).max() < AMP_TOL


def test_full_size_L20_trajectories_vs_c_oracle(ctx, disorder):
    """BASELINE config C2 shape (n = 21, complex128): amplitude parity on whole 2^21 states."""
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", 20, 0.97, hs, phis, 3, 10, True))
    oc, na, _ = O.compact_ops(RC.ops_of(circ), 31)
    prog = compile_circuit(circ, RC.noise_model(0.05))
    assert prog.n == 21
    psi, batch = _evolve_true(ctx, prog, 3, 17, 1234)
    onoise = O.PauliNoise.depolarizing(0.05)
    for i, tr in enumerate((17, 18, 19)):
        ref = CO.run_trajectory(oc, na, onoise, 1234, tr)
        assert np.abs(psi[i] - ref).max() < AMP_TOL


def test_readout_kernels(ctx, disorder):
    """dtc_probs / dtc_expect_z / dtc_sample_states against numpy on the same states."""
    from dtcsim import backend
    hs, phis = disorder[20][0][3][:13], disorder[20][1][3][:12]
    circ = RC.transpiled(RC.qc_body("vacuum", 13, 0.9, hs, phis, 2, 6, False))
    prog = compile_circuit(circ, RC.noise_model(0.3))
    batch = backend.evolve(ctx, prog, 16, 0, 3)
    pz = batch.expect_z().cpu().numpy()
    q3 = [prog.measures[0][0], 2, 7]
    p3 = batch.probs(q3).cpu().numpy()
    p1 = batch.probs([prog.measures[0][0]]).cpu().numpy()
    psi = batch.materialize().cpu().numpy()
    pr = np.abs(psi) ** 2
    idx = np.arange(1 << prog.n)
    for q in range(prog.n):
        assert np.abs(pz[:, q] - pr @ (1.0 - 2.0 * ((idx >> q) & 1))).max() < 1e-12
    col = sum(((idx >> q) & 1) << i for i, q in enumerate(q3))
    for r in range(16):
        assert np.abs(p3[r] - np.bincount(col, weights=pr[r], minlength=8)).max() < 1e-12
    assert np.abs(p1.sum(1) - 1).max() < 1e-12 and np.abs(p1[:, 1] - (1 - pz[:, q3[0]]) / 2).max() < 1e-12


def test_sample_states_distribution(ctx):
    from dtcsim import backend
    c = dtcsim.QuantumCircuit(13, 13)
    c.h(0)
    for q in range(12):
        c.cx(q, q + 1)
    c.measure_all()
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.pauli_error([("X", 0.5), ("I", 0.5)]), ["h"])
    prog = compile_circuit(c, nm)
    batch = backend.evolve(ctx, prog, 4000, 0, 11)
    idx = batch.sample_states(11).cpu().numpy()
    assert set(np.unique(idx)) <= {0, (1 << 13) - 1}            # GHZ: all zeros or all ones
    frac = (idx != 0).mean()
    assert abs(frac - 0.5) < 4 * 0.5 / np.sqrt(4000)


# ----------------------------------------------------------------------------------- density matrix
def test_density_matrix_known_answers(disorder, known):
    """Reference config C1 through run(): exact DM expectation vs SURVEY 8c values (1e-8)."""
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    for t, (f, e) in known["survey_8c"]["L4_g0.84_p0.05"].items():
        for echo, want in ((False, f), (True, e)):
            circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, int(t), 2, echo), backend=sim)
            res = sim.run(circ, shots=1024, seed_simulator=5).result()
            assert res.data()["method"] == "density_matrix"
            assert abs(res.expectation_z()[0] - want) < DM_TOL


def test_density_matrix_elementwise(ctx, disorder):
    from dtcsim import backend
    hs, phis = disorder[20][0][0][:6], disorder[20][1][0][:5]
    circ = RC.transpiled(RC.qc_body("neel", 6, 0.9, hs, phis, 3, 3, True, "yx"))
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.pauli_error([("X", 0.05), ("Y", 0.02), ("Z", 0.1), ("I", 0.83)]), ["u2", "u3"])
    prog = compile_circuit(circ, nm, want_dm=True)
    rho = backend.run_density_matrix(ctx, prog).cpu().numpy()
    d = 1 << prog.n
    rho = PI.dm_to_circuit_order(rho.reshape(d, d).T, prog)
    oc, na, _ = O.compact_ops(RC.ops_of(circ), 31)
    ref = O.run_density_matrix(oc, na, O.PauliNoise({"u2": (0.05, 0.02, 0.1), "u3": (0.05, 0.02, 0.1)}))
    assert np.abs(rho - ref).max() < 1e-12
    assert abs(np.trace(rho) - 1) < 1e-12


def test_density_matrix_L12_config(disorder):
    """BASELINE config C3: L = 12 ancilla-free chain (rho has 2^24 entries), <Z_6> vs exact light cone."""
    hs, phis = disorder[20][0][0][:12], disorder[20][1][0][:11]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), method="density_matrix")
    for t in (1, 2, 3):
        c = dtcsim.QuantumCircuit(12, 1)
        for _ in range(t):
            for i in range(12):
                c.rx(np.pi * 0.97, i)
            for i in range(0, 11, 2):
                c.rzz(phis[i], i, i + 1)
            for i in range(1, 11, 2):
                c.rzz(phis[i], i, i + 1)
            for i in range(12):
                c.rz(hs[i], i)
        c.measure(6, 0)
        low = dtcsim.lower_level0(c)
        res = sim.run(low, shots=100, seed_simulator=1).result()
        want = O.lightcone_zq(12, 0.97, hs, phis, t, 6, 0.05)
        assert abs(res.expectation_z()[0] - want) < DM_TOL
    # 20 periods: trace preserved and |<Z>| <= 1
    c = dtcsim.QuantumCircuit(12, 1)
    for _ in range(20):
        for i in range(12):
            c.rx(np.pi * 0.97, i)
        for i in range(0, 11, 2):
            c.rzz(phis[i], i, i + 1)
        for i in range(1, 11, 2):
            c.rzz(phis[i], i, i + 1)
        for i in range(12):
            c.rz(hs[i], i)
    c.measure(6, 0)
    res = sim.run(dtcsim.lower_level0(c), shots=100, seed_simulator=1).result()
    pr = res.data()["probabilities"]
    assert abs(sum(pr.values()) - 1) < 1e-10 and abs(res.expectation_z()[0]) < 1


# ----------------------------------------------------------------------------------- run() / counts
def test_counts_bit_exact_vs_oracle(disorder):
    """Same Philox contract on both sides: get_counts() equals the oracle's counts exactly."""
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    onoise = O.PauliNoise.depolarizing(0.05)
    for t, echo, shots in ((1, False, 1024), (3, True, 1024), (2, False, 16), (4, True, 30)):
        circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, t, 2, echo), backend=sim)
        got = sim.run(circ, shots=shots, seed_simulator=1234).result().get_counts(circ)
        want, info = O.run_counts(RC.ops_of(circ), 31, 1, shots=shots, noise=onoise, seed=1234)
        assert got == want, (t, echo, shots, got, want, info["method"])
    # ideal circuit: one statevector, `shots` samples
    sim0 = dtcsim.AerSimulator()
    circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, 3, 2, False))
    got = sim0.run(circ, shots=500, seed_simulator=7).result().get_counts()
    want, _ = O.run_counts(RC.ops_of(circ), 31, 1, shots=500, noise=None, seed=7)
    assert got == want
    # trajectories at n = 12 (shots < 2^n)
    hs, phis = disorder[20][0][0][:11], disorder[20][1][0][:10]
    circ = RC.transpiled(RC.qc_body("vacuum", 11, 0.97, hs, phis, 2, 5, False))
    got = sim.run(circ, shots=200, seed_simulator=99).result().get_counts()
    want, info = O.run_counts(RC.ops_of(circ), 31, 1, shots=200, noise=onoise, seed=99)
    assert info["method"] == "statevector" and got == want


def test_multi_clbit_counts_dtc_qasm_shape(disorder):
    """dtc_qasm.py circuit shape: L qubits, measure all; per-qubit <Z_i> from counts (dtc_qasm.py:145)."""
    hs, phis = disorder[20][0][0][:6], disorder[20][1][0][:5]
    ops, n, nc = C.dtc_qasm_gates("1", 6, 0.94, hs, phis, 3)
    circ = dtcsim.QuantumCircuit(6, 6)
    for nm_, qs, ps, cs in ops:
        circ._add(nm_, qs, ps, cs)
    sim = dtcsim.AerSimulator()
    res = sim.run(circ, shots=4096, seed_simulator=3).result()
    counts = res.get_counts()
    want, info = O.run_counts(ops, 6, 6, shots=4096, noise=None, seed=3)
    assert counts == want
    ez = dtcsim.backend.compute_z_expectation(counts, 6)
    psi = O.run_statevector(ops, 6)
    idx = np.arange(64)
    for q in range(6):
        exact = float(np.sum(np.abs(psi) ** 2 * (1 - 2 * ((idx >> q) & 1))))
        assert abs(ez[q] - exact) < 4 / np.sqrt(4096)
        assert abs(res.expectation_z()[q] - exact) < 1e-10


def test_L20_trajectory_statistics(disorder):
    """C2-shaped point: 1024 Pauli trajectories at n = 21 within 4 sigma of the exact value and of the
    reference's own committed Aer output (autocorr_data_L20_polarization/*polx*.csv)."""
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    df = golden_csv("ref_L20_pol_x.csv")
    for t, echo in ((2, False), (1, True)):
        circ = RC.transpiled(RC.qc_body("vacuum", 20, 0.97, hs, phis, t, 10, echo), backend=sim)
        res = sim.run(circ, shots=1024, seed_simulator=1234).result()
        assert res.data()["method"] == "statevector" and res.data()["n_qubits"] == 21
        counts = res.get_counts(circ)
        ez = dtcsim.backend.compute_z_expectation(counts, 1)[0]
        exact = 0.95 ** 6 * O.lightcone_zq(20, 0.97, hs, phis, t, 10, 0.05, echo=echo)
        sig = np.sqrt((1 - exact ** 2) / 1024)
        zf = family_z(6)                                              # 2 circuits x 3 comparisons in this test
        assert abs(ez - exact) < zf * sig
        assert abs(res.expectation_z()[0] - exact) < zf * sig         # trajectory mean of exact probabilities
        ref = df["av_autocorr_echo" if echo else "av_autocorr"][t]
        assert abs(ez - ref) < zf * np.sqrt(2) * sig                  # two independent 1024-shot estimates


def test_L20_circular_and_controlled_g_through_run(disorder):
    """The circular-polarisation and time-dependent-g anchors of SURVEY 8c through run() at full size (n = 21, 1024
    trajectories): circ-pol.py:162-174 (per-step rx/ry angles, two noisy u3 per qubit per period) and g-opt.py:530-545
    (step k uses g_history_inst1[k]); against the exact light-cone value and the reference's committed Aer output."""
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    dfc = golden_csv("ref_L20_circ_circular_left.csv")
    dfg = golden_csv("ref_L20_controlled_iter5.csv")
    gh = [float(x) for x in dfg["g_history_inst1"]]
    cases = []
    # circular_left, forward t = 2 and echo t = 1
    for t, echo in ((2, False), (1, True)):
        ops, _, _ = C.autocorr_gates("vacuum", 20, 0.97, hs, phis, t, 10, echo, polarization="circular_left",
                                     circular_frequency=1.0)
        pf = lambda step: C.uf_gates(20, 0.97, phis, hs, "circular_left", time_step=step, circular_frequency=1.0)
        exact = 0.95 ** 6 * O.lightcone_zq(20, 0.97, hs, phis, t, 10, 0.05, echo=echo, period_fn=pf)
        cases.append((ops, exact, dfc["av_autocorr_echo" if echo else "av_autocorr"][t]))
    # time-dependent g (row i <-> i+1 periods), forward 3 periods and echo 2 periods
    for t, echo in ((3, False), (2, True)):
        ops, _, _ = C.autocorr_gates("vacuum", 20, None, hs, phis, t, 10, echo, g_values=gh)
        pf = lambda step: C.uf_gates(20, gh[step], phis, hs, "x")
        exact = 0.95 ** 6 * O.lightcone_zq(20, None, hs, phis, t, 10, 0.05, echo=echo, period_fn=pf)
        cases.append((ops, exact, dfg["echo_adaptive_inst1" if echo else "forward_adaptive_inst1"][t - 1]))
    zf = family_z(2 * len(cases))
    for i, (ops, exact, ref) in enumerate(cases):
        circ = dtcsim.QuantumCircuit(31, 1)
        for nm_, qs, ps, cs in C.lower_level0(ops, C.SNAKE_LAYOUT):
            circ._add(nm_, qs, ps, cs)
        res = sim.run(circ, shots=1024, seed_simulator=4321 + i).result()
        assert res.data()["method"] == "statevector" and res.data()["n_qubits"] == 21
        ez = dtcsim.backend.compute_z_expectation(res.get_counts(), 1)[0]
        sig = np.sqrt((1 - exact ** 2) / 1024)
        assert abs(ez - exact) < zf * sig, (i, ez, exact)
        assert abs(ez - ref) < zf * np.sqrt(2) * sig, (i, ez, ref)


def test_xy_cycle_amplitudes_vs_oracle(ctx, disorder):
    """xy-cycle.py:141-157: the kick polarisation alternates x / y every 5 steps.  t = 7 crosses the switch; noisy
    trajectory amplitudes (forward and echo) against the oracle at 1e-10."""
    L = 11
    hs, phis = disorder[20][0][2][:L], disorder[20][1][2][:L - 1]
    sched = lambda step: "x" if (step // 5) % 2 == 0 else "y"
    for echo in (False, True):
        ops, _, _ = C.autocorr_gates("vacuum", L, 0.97, hs, phis, 7, L // 2, echo, pol_schedule=sched)
        low = C.lower_level0(ops, C.SNAKE_LAYOUT)
        circ = dtcsim.QuantumCircuit(31, 1)
        for nm_, qs, ps, cs in low:
            circ._add(nm_, qs, ps, cs)
        oc, na, _ = O.compact_ops(low, 31)
        prog = compile_circuit(circ, RC.noise_model(0.05))
        psi, _ = _evolve_true(ctx, prog, 5, 11, 2024, engine=2)
        ref = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(0.05), 2024, np.arange(11, 16))
        assert np.abs(psi - ref).max() < AMP_TOL


def test_full_size_L20_counts_equal_c_oracle(disorder):
    """get_counts() of the factorised n = 20 pipeline (k_frames -> k_tile_stream -> fused read-out -> k_readout_small ->
    k_sample_rows) == the gate-by-gate C oracle on the full n = 21 register, trajectory by trajectory, at the size of
    BASELINE config C2 (64 trajectories, forward and echo)."""
    from oracle import philox
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    onoise = O.PauliNoise.depolarizing(0.05)
    shots = 64
    for t, echo, seed in ((3, False, 1234), (2, True, 77)):
        circ = RC.transpiled(RC.qc_body("vacuum", 20, 0.97, hs, phis, t, 10, echo))
        res = sim.run(circ, shots=shots, seed_simulator=seed).result()
        assert res.data()["register_qubits"] == 20 and res.data()["n_qubits"] == 21
        oc, na, _ = O.compact_ops(RC.ops_of(circ), 31)
        (mq, _c), = O.measured_map(oc)
        buf = np.empty(1 << na, dtype=np.complex128)
        vals, p1s = [], []
        for tr in range(shots):
            psi = CO.run_trajectory(oc, na, onoise, seed, tr, out=buf)
            p1 = CO.prob1(psi, na, mq)
            u = philox.uniform(seed, 0, philox.STREAM_MEASURE, np.array([tr], dtype=np.uint64))[0]
            vals.append(int(O.sample_outcome(np.cumsum([1.0 - p1, p1]), u)))
            p1s.append(p1)
        want = O.counts_dict(vals, 1)
        assert res.get_counts() == want, (t, echo, res.get_counts(), want)
        assert abs(res.expectation_z()[0] - (1 - 2 * np.mean(p1s))) < 1e-10


def test_density_matrix_every_probe_qubit(disorder):
    """run() on the exact density-matrix path for EVERY probe site q (and t = 0): the measured bit must be taken from the
    program that built rho (its bit order differs from the factorised trajectory program's)."""
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    onoise = O.PauliNoise.depolarizing(0.05)
    for L, row in ((4, 4), (6, 20)):
        hs, phis = disorder[row][0][0][:L], disorder[row][1][0][:L - 1]
        for q in range(L):
            for t, echo in ((0, False), (3, False), (2, True)):
                circ = RC.transpiled(RC.qc_body("vacuum", L, 0.84, hs, phis, t, q, echo), backend=sim)
                res = sim.run(circ, shots=1024, seed_simulator=11).result()
                assert res.data()["method"] == "density_matrix"
                want, info = O.run_counts(RC.ops_of(circ), 31, 1, shots=1024, noise=onoise, seed=11)
                pr = info["probabilities"]
                assert abs(res.expectation_z()[0] - (pr[0] - pr[1])) < DM_TOL, (L, q, t, echo)
                assert res.get_counts() == want, (L, q, t, echo)


def test_noiseless_echo_returns_to_start(disorder):
    """Size-independent property at the full configuration: U^-t U^t = 1, so P(anc = 0) = 1."""
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    sim = dtcsim.AerSimulator()
    circ = RC.transpiled(RC.qc_body("vacuum", 20, 0.97, hs, phis, 7, 10, True))
    res = sim.run(circ, shots=64, seed_simulator=1).result()
    assert abs(res.expectation_z()[0] - 1.0) < 1e-10
    assert res.get_counts() == {"0": 64}


def test_norm_preserved_full_batch(ctx, disorder):
    from dtcsim import backend
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", 20, 0.97, hs, phis, 6, 10, True))
    prog = compile_circuit(circ, RC.noise_model(0.05))
    batch = backend.evolve(ctx, prog, 64, 0, 5)
    p = batch.probs([prog.measures[0][0]]).cpu().numpy()
    assert np.abs(p.sum(1) - 1).max() < 1e-11


def test_error_paths(ctx):
    sim = dtcsim.AerSimulator()
    c = dtcsim.QuantumCircuit(2, 1)
    c.h(0)
    with pytest.raises(ValueError):
        sim.run(c, shots=10).result()            # nothing measured
    with pytest.raises(ValueError):
        sim.run(c, shots=0)
    from dtcsim import capi
    import ctypes
    h = ctypes.c_void_p()
    assert capi.load().dtc_program_create(0, 1, ctypes.byref(h)) == -1
    assert b"n_qubits" in capi.load().dtc_last_error()


# ----------------------------------------------------------------------------------- read-out factorisation
def test_rdm_kernel(ctx, disorder):
    from dtcsim import backend
    hs, phis = disorder[20][0][3][:13], disorder[20][1][3][:12]
    circ = RC.transpiled(RC.qc_body("vacuum", 13, 0.9, hs, phis, 2, 6, False))
    prog = compile_circuit(circ, RC.noise_model(0.3))
    batch = backend.evolve(ctx, prog, 8, 0, 3)
    r2 = batch.rdm([5, 2]).cpu().numpy()
    r1 = batch.rdm([11]).cpu().numpy()
    r0 = batch.rdm([]).cpu().numpy()
    psi = batch.state.view(8, -1).cpu().numpy()
    from test_readout_factorisation import _rdm
    assert np.abs(r2 - _rdm(psi, prog.n, [5, 2])).max() < 1e-12
    assert np.abs(r1 - _rdm(psi, prog.n, [11])).max() < 1e-12
    assert np.abs(r0[:, 0, 0] - 1).max() < 1e-12


def test_factorised_run_equals_full_run(disorder):
    """run() with the ancilla factorised out (default) == run() on the full register == oracle counts."""
    hs, phis = disorder[20][0][0][:12], disorder[20][1][0][:11]
    nm = RC.noise_model(0.05)
    fast = dtcsim.AerSimulator(noise_model=nm)
    full = dtcsim.AerSimulator(noise_model=nm, optimize=False)
    for t, echo, state in ((0, False, "vacuum"), (2, False, "neel"), (3, True, "vacuum")):
        circ = RC.transpiled(RC.qc_body(state, 12, 0.97, hs, phis, t, 6, echo))
        r1 = fast.run(circ, shots=300, seed_simulator=42).result()
        r2 = full.run(circ, shots=300, seed_simulator=42).result()
        assert r1.data()["register_qubits"] == (12 if t else 1) and r2.data()["register_qubits"] == (13 if t else 2)
        assert r1.get_counts() == r2.get_counts()
        assert abs(r1.expectation_z()[0] - r2.expectation_z()[0]) < 1e-12
        want, _ = O.run_counts(RC.ops_of(circ), 31, 1, shots=300, noise=O.PauliNoise.depolarizing(0.05), seed=42)
        assert r1.get_counts() == want


def test_trajectory_ranges_concatenate(disorder):
    """Multi-GPU contract (dist.ShardedSampler): disjoint trajectory ranges reproduce the single-GPU run."""
    hs, phis = disorder[20][0][0][:12], disorder[20][1][0][:11]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05))
    circ = RC.transpiled(RC.qc_body("vacuum", 12, 0.97, hs, phis, 3, 6, True))
    whole = sim.sample_trajectories(circ, 0, 200, 9)
    parts = np.concatenate([sim.sample_trajectories(circ, 0, 77, 9), sim.sample_trajectories(circ, 77, 200, 9)])
    assert np.array_equal(whole, parts)
    counts = sim.run(circ, shots=200, seed_simulator=9).result().get_counts()
    assert counts == {k: v for k, v in (("0", int((whole == 0).sum())), ("1", int((whole == 1).sum()))) if v}
    from dtcsim import dist as D
    sampler = D.ShardedSampler(sim, 0, 1)
    assert sampler.run_counts(circ, shots=200, seed_simulator=9) == counts


# ---- TMA-fed streaming engine (k_tile_stream)
@pytest.mark.parametrize("L,t,echo,pol,state,ntraj", [
    (12, 2, False, "x", "vacuum", 200),    # n = 12: more tiles than SMs, one contiguous tile per state
    (13, 3, True, "y", "vacuum", 5),       # mode A + mode B (64 B runs, g = 3), fewer tiles than CTAs x stages
    (16, 2, True, "xy", "neel", 21),       # ragged: 21 * 16 tiles over 148 CTAs
    (20, 3, True, "x", "vacuum", 3),       # config-C2 shape: 768 tiles, >= 5 per CTA (stage ring wraps)
])
def test_stream_engine_equals_register_engine(ctx, disorder, L, t, echo, pol, state, ntraj):
    """Same program through k_tile_stream and through k_tile_pass: psi' within 1e-13, frames identical."""
    from dtcsim import backend, capi
    hs, phis = disorder[20][0][1][:L], disorder[20][1][1][:L - 1]
    circ = RC.transpiled(RC.qc_body(state, L, 0.97, hs, phis, t, L // 2, echo, pol))
    prog = compile_circuit(circ, RC.noise_model(0.05), optimize=True)
    h = capi.ProgramHandle(prog, 0)
    assert h.num_stream_passes == h.num_passes > 0
    try:
        # init_index != 0 also checks the first pass generating the basis state itself (no memset, no read)
        init = ((0x2A5 << (L - 10)) | 3) if L >= 13 else 0
        capi.set_stream_engine(False)
        a = backend.evolve(ctx, prog, ntraj, 7, 99, handle=h, init_index=init)
        sa, fa = a.state.clone(), a.frames_host()
        capi.set_stream_engine(True)
        garbage = ctx.empty(ntraj << prog.n_main, a.state.dtype).fill_(float("nan"))   # the input buffer is never read
        b = backend.evolve(ctx, prog, ntraj, 7, 99, handle=h, init_index=init, state=garbage)
        sb, fb = b.state, b.frames_host()
    finally:
        capi.set_stream_engine(None)
    assert all(np.array_equal(x, y) for x, y in zip(fa, fb))
    assert float((sa - sb).abs().max()) < 1e-13
    assert abs(float((sb.abs() ** 2).sum()) / ntraj - 1.0) < 1e-9


def test_stream_engine_vs_oracle_full_register(ctx, disorder):
    """Ancilla kept in the register: streaming where eligible, amplitudes against the numpy oracle."""
    from dtcsim import capi
    L = 14
    hs, phis = disorder[20][0][2][:L], disorder[20][1][2][:L - 1]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, 2, L // 2, True, "x"))
    oc, na, _ = O.compact_ops(RC.ops_of(circ), circ.num_qubits)
    prog = compile_circuit(circ, RC.noise_model(0.05))
    h = capi.ProgramHandle(prog, 0)
    assert h.num_stream_passes > 0
    psi, _ = _evolve_true(ctx, prog, 4, 11, 5, engine=2)
    ref = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(0.05), 5, np.arange(11, 15))
    assert np.abs(psi - ref).max() < AMP_TOL


@pytest.mark.parametrize("L,t,echo,ntraj", [(12, 2, True, 40), (16, 3, False, 9), (20, 2, True, 3)])
def test_fused_readout_rdm_equals_rdm_kernel(ctx, disorder, L, t, echo, ntraj):
    """Last pass reducing the read-out qubit's density matrix (state not stored) == dtc_rdm on the stored state,
    and the outcome probabilities built from either agree."""
    from dtcsim import backend, capi
    hs, phis = disorder[20][0][3][:L], disorder[20][1][3][:L - 1]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, t, L // 2, echo, "x"))
    prog = compile_circuit(circ, RC.noise_model(0.05), optimize=True)
    h = capi.ProgramHandle(prog, 0)
    a = backend.evolve(ctx, prog, ntraj, 3, 17, handle=h)
    assert a.fused_rdm is None
    rdm_ref = a.rdm(prog.small["reg_bits"]).clone()
    pr_ref = a.outcome_probs().clone()
    b = backend.evolve(ctx, prog, ntraj, 3, 17, handle=h, fused_rdm=True)
    if h.num_stream_passes == h.num_passes:
        assert b.fused_rdm is not None          # every pass streams, the read-out qubit sits in the last tile
    if b.fused_rdm is not None:
        assert float((b.fused_rdm - rdm_ref).abs().max()) < 1e-12
    assert float((b.outcome_probs() - pr_ref).abs().max()) < 1e-12
    assert abs(float(pr_ref.sum()) / ntraj - 1.0) < 1e-10


def test_run_list_is_pipelined_and_equals_single_runs(disorder):
    """run([c0, c1, c2]) (circuit i+1 enqueued before circuit i is read back) == three run() calls with seeds s, s+1, s+2."""
    hs, phis = disorder[20][0][0][:13], disorder[20][1][0][:12]
    circs = [RC.transpiled(RC.qc_body("vacuum", 13, 0.97, hs, phis, t, 6, echo)) for t, echo in ((1, False), (3, True), (2, False))]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05))
    res = sim.run(circs, shots=300, seed_simulator=40).result()
    for i, c in enumerate(circs):
        one = sim.run(c, shots=300, seed_simulator=40 + i).result()
        assert res.get_counts(c) == one.get_counts(c) == res.get_counts()[i]
        assert abs(res.expectation_z(i)[0] - one.expectation_z()[0]) < 1e-12


def test_wide_measurement_ideal_and_noisy(disorder):
    """dtc_qasm.py shape at L = 14 (> 12 measured qubits): shots are basis-state samples; exact per-qubit <Z> comes from
    one dtc_expect_z pass.  Ideal: against the oracle statevector; noisy: against the oracle's trajectories (same Philox ids)."""
    L = 14
    hs, phis = disorder[20][0][0][:L], disorder[20][1][0][:L - 1]
    ops, n, nc = C.dtc_qasm_gates("1", L, 0.94, hs, phis, 2)
    circ = dtcsim.QuantumCircuit(L, L)
    for nm_, qs, ps, cs in ops:
        circ._add(nm_, qs, ps, cs)
    psi = O.run_statevector(ops, L)
    idx = np.arange(1 << L)
    exact = np.array([np.sum(np.abs(psi) ** 2 * (1 - 2 * ((idx >> q) & 1))) for q in range(L)])
    shots = 4096
    res = dtcsim.AerSimulator().run(circ, shots=shots, seed_simulator=5).result()
    counts = res.get_counts()
    assert sum(counts.values()) == shots and all(len(k) == L for k in counts)
    assert np.abs(np.array(res.expectation_z()) - exact).max() < 1e-10
    ez = np.array(dtcsim.backend.compute_z_expectation(counts, L))
    assert np.abs(ez - exact).max() < 4.5 / np.sqrt(shots)
    # noisy: 96 trajectories, every qubit measured
    nmod = dtcsim.NoiseModel()
    nmod.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.2, 1), ["rx", "x"])
    ntraj = 96
    resn = dtcsim.AerSimulator(noise_model=nmod).run(circ, shots=ntraj, seed_simulator=8).result()
    oc, na, _ = O.compact_ops(ops, L)
    psin = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(0.2, names=("rx", "x")), 8, np.arange(ntraj))
    want = np.array([np.mean(np.sum(np.abs(psin) ** 2 * (1 - 2 * ((idx >> q) & 1)), axis=1)) for q in range(L)])
    assert np.abs(np.array(resn.expectation_z()) - want).max() < 1e-10
    assert sum(resn.get_counts().values()) == ntraj


@pytest.mark.parametrize("n,ntraj", [(17, 5), (19, 2)])
def test_stream_engine_mode_c_high_stride_groups(ctx, n, ntraj):
    """Groups starting at the high-stride bit (lowered to 10 here; 15 in production) get five qubits and tiles of 32 runs of
    2 KB (4-D tensor map, one in-place phase): against the register-fed kernel and the oracle."""
    from dtcsim import backend, capi
    rng = np.random.default_rng(50 + n)
    c = dtcsim.QuantumCircuit(n, 0)
    for layer in range(3):
        for q in range(n):
            c.rx(rng.uniform(-3, 3), q)
        for q in range(n - 1):
            c.rzz(rng.uniform(-3, 3), q, q + 1)
        for q in range(n):
            c.rz(rng.uniform(-3, 3), q)
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.3, 1), ["rx"])
    prog = compile_circuit(c, nm, reorder=False)
    capi.set_high_stride_bit(10)
    try:
        h = capi.ProgramHandle(prog, 0)
        assert h.num_stream_passes == h.num_passes
        capi.set_stream_engine(False)
        a = backend.evolve(ctx, prog, ntraj, 2, 9, handle=h)
        sa = a.state.clone()
        capi.set_stream_engine(True)
        b = backend.evolve(ctx, prog, ntraj, 2, 9, handle=h)
        assert float((sa - b.state).abs().max()) < 1e-13
        psi = PI.to_circuit_order(b.materialize().cpu().numpy(), prog)
    finally:
        capi.set_stream_engine(None)
        capi.set_high_stride_bit(15)
    oc, na, _ = O.compact_ops(RC.ops_of(c), n)
    ref = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(0.3, names=("rx",)), 9, np.arange(2, 2 + ntraj))
    assert np.abs(psi - ref).max() < AMP_TOL


@pytest.mark.parametrize("seed", range(10))
def test_engines_agree_on_random_circuits(ctx, seed):
    """Fuzz: general gate sets, odd tile shapes, mixed eligible / ineligible passes -- k_tile_stream wherever the planner
    allows it against k_tile_pass everywhere, same frames, psi' within 1e-12."""
    from dtcsim import backend, capi
    rng = np.random.default_rng(900 + seed)
    n = 12 + seed % 6
    circ = _random_circuit(rng, n, 50 + 5 * seed)
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.2, 1), ["u1", "u2", "u3", "h"])
    prog = compile_circuit(circ, nm)
    h = capi.ProgramHandle(prog, 0, capi.ENGINE_TILE)
    ntraj = 1 + seed % 4
    try:
        capi.set_stream_engine(False)
        a = backend.evolve(ctx, prog, ntraj, 1, seed, handle=h)
        sa, fa = a.state.clone(), a.frames_host()
        capi.set_stream_engine(True)
        b = backend.evolve(ctx, prog, ntraj, 1, seed, handle=h)
        fb = b.frames_host()
    finally:
        capi.set_stream_engine(None)
    assert all(np.array_equal(x, y) for x, y in zip(fa, fb))
    assert float((sa - b.state).abs().max()) < 1e-12


# ---- sharded statevector (config C5) through the C ABI
def _dtc_chain_circuit(L, t, rng, echo=False):
    hs = rng.random(L) * 2 * np.pi - np.pi                      # generate_disorder.py:16-18
    phis = rng.random(L - 1) * np.pi - 1.5 * np.pi
    ops, _, _ = C.dtc_qasm_gates("1", L, 0.97, hs, phis, t)
    if echo:
        body = [o for o in ops if o[0] != "measure"]
        ops = body + C.inverse_gates([o for o in body if o[0] != "x"]) + [o for o in ops if o[0] == "measure"]
    low = C.lower_level0(ops)
    circ = dtcsim.QuantumCircuit(L, L)
    for nm_, qs, ps, cs in low:
        circ._add(nm_, qs, ps, cs)
    return circ, low


@pytest.mark.parametrize("world,L,high_bit,ce_quarters", [(1, 14, 15, 0), (2, 16, 15, 0), (4, 17, 9, 0), (2, 18, 10, 0),
                                                         (2, 18, 15, 2), (4, 19, 15, 1)])
def test_sharded_engine_vs_oracle(world, L, high_bit, ce_quarters):
    """CudaShardEngine + ShardedStatevector (rank bits in the diagonal phases, one-sweep top-group segments, slice
    programs fused with the exchange, qubit permutation kept) against the oracle: a noisy trajectory of the dtc_qasm.py
    circuit shape.  world > 1: the ranks are emulated by threads on this one GPU (sharded.ThreadFabric)."""
    import threading
    from dtcsim import capi, sharded
    rng = np.random.default_rng(34 + L)
    circ, low = _dtc_chain_circuit(L, 3, rng)
    noise = RC.noise_model(0.2)
    g = int(np.log2(world))
    fabric = sharded.ThreadFabric(world) if world > 1 else None
    results, errors = [None] * world, []

    def work(rank):
        try:
            eng = sharded.CudaShardEngine(L, L - g, rank, world, 0, fabric=fabric)
            allred = (lambda a: fabric.all_reduce(rank, a)) if fabric else None
            sv = sharded.ShardedStatevector(L, rank, world, eng, all_reduce=allred)
            r = sv.run(circ, noise, seed=11, trajectory=5)
            results[rank] = (r, dict(sv.stats), eng.sliced_exchanges, eng.fused_stores)
            eng.close()
        except Exception as exc:                      # surfaces in the main thread; peers fail on the broken barrier
            errors.append(exc)
            if fabric:
                fabric.bar.abort()

    capi.set_high_stride_bit(high_bit)
    saved = sharded.CudaShardEngine.CE_QUARTERS
    sharded.CudaShardEngine.CE_QUARTERS = ce_quarters      # part of every slice copied instead of stored by the last sweep
    try:
        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
    finally:
        capi.set_high_stride_bit(15)
        sharded.CudaShardEngine.CE_QUARTERS = saved
    assert not errors, errors
    oc, na, _ = O.compact_ops(low, L)
    psi = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(0.2), 11, [5])[0]
    idx = np.arange(1 << L)
    want = np.array([np.sum(np.abs(psi) ** 2 * (1 - 2 * ((idx >> q) & 1))) for q in range(L)])
    for r, stats, n_sliced, n_fused in results:
        assert abs(r["norm"] - 1) < 1e-11
        assert np.abs(np.array(r["expect_z"]) - want).max() < AMP_TOL
        if world > 1:
            assert stats["exchanges"] == n_sliced > 0 and stats["exchanges"] <= stats["layers"]
            assert n_fused > 0          # the last sweep of the slice programs stored straight into the receivers' buffers


# ---- resident execution (k_tile_resident): all sweeps of a circuit in one persistent launch over L2-resident groups
@pytest.mark.parametrize("L,t,echo,ntraj,resident_mb,high_bit", [
    (12, 3, True, 37, 64, 15),        # one tile per state, everything in one group
    (14, 2, False, 11, 1, 15),        # 4 tiles per state, groups of 4 slots, ragged last group
    (16, 3, True, 9, 1, 15),          # groups of ONE slot: every item waits for the previous pass of its own slot
    (16, 2, True, 7, 2, 9),           # high-stride groups: 2 KB-run tiles (second tensor map) in the same launch
    (20, 2, True, 6, 64, 15),         # the C2 register: groups of 4 states of 16 MiB
    (20, 3, False, 3, 16, 15),        # ... and of one state
])
def test_resident_execution_equals_streamed(ctx, disorder, L, t, echo, ntraj, resident_mb, high_bit):
    """Fused read-out density matrices and frames of a resident run == those of the sweep-per-launch run (same kernels'
    phases, different control structure: work order, completion counters, state slots reused by consecutive groups)."""
    from dtcsim import backend, capi
    hs, phis = disorder[20][0][1][:L], disorder[20][1][1][:L - 1]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, t, L // 2, echo))
    capi.set_high_stride_bit(high_bit)
    try:
        prog = compile_circuit(circ, RC.noise_model(0.1), optimize=True)
        assert prog.small is not None and prog.n_main == L
        capi.RESIDENT = False
        ref = backend.evolve(ctx, prog, ntraj, 5, 99, fused_rdm=True)
        assert not ref.resident
        rdm_ref = ref.readout_rdm().clone()
        probs_ref = ref.outcome_probs().clone()
        capi.RESIDENT = True
        capi.set_resident_bytes(resident_mb << 20)
        h = capi.ProgramHandle(prog, 0)
        got = backend.evolve(ctx, prog, ntraj, 5, 99, handle=h, fused_rdm=True)
        if ref.fused_rdm is None:
            pytest.skip("last pass of this schedule is not fusable: resident execution does not apply")
        assert got.resident and h.last_run_info() == (True, 2)
        assert float((got.readout_rdm() - rdm_ref).abs().max()) < 1e-12
        assert float((got.outcome_probs() - probs_ref).abs().max()) < 1e-12
        assert all(np.array_equal(a, b) for a, b in zip(got.frames_host(), ref.frames_host()))
        # a second run on the same handle and scratch (counters and slots are reset per run)
        got2 = backend.evolve(ctx, prog, ntraj, 5, 99, handle=h, state=got.state, fused_rdm=True)
        assert float((got2.readout_rdm() - rdm_ref).abs().max()) < 1e-12
        h.close()
    finally:
        capi.RESIDENT = False
        capi.set_resident_bytes(64 << 20)
        capi.set_high_stride_bit(15)


def test_run_sweep_equals_per_point_oracle(disorder):
    """sweeps.run_sweep (g x polarisation x echo x instance x t in one call) returns, point by point, what the reference's
    loop returns: <Z_ancilla> from the counts of the point's circuit with seed = seed + global point index -- checked
    against the oracle's counts of the same circuits (bit-identical under the shared Philox contract)."""
    L = 6
    hs, phis = disorder[20][0][:2, :L], disorder[20][1][:2, :L - 1]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    g_list, pols, t_values, echoes = [0.84, 0.97], ("x", "yx"), [0, 1, 3], (False, True)
    res = dtcsim.run_sweep(sim, L, g_list, hs, phis, t_values, echoes, pols, shots=64, seed_simulator=500, chunk=7)
    assert res["autocorr"].shape == (2, 2, 2, 2, 3) and res["points"] == 48
    assert res["periods"] == 64 * 2 * 2 * 2 * (sum(t_values) + 2 * sum(t_values))
    onoise = O.PauliNoise.depolarizing(0.05)
    from dtcsim import sweeps
    pts = sweeps.sweep_points(g_list, pols, range(2), t_values, echoes)
    for k in (0, 5, 17, 30, 47):
        gi, pi, ei, ii, ti = pts[k]
        ops, _, _ = C.autocorr_gates("vacuum", L, g_list[gi], hs[ii], phis[ii], t_values[ti], L // 2, echoes[ei],
                                     polarization=pols[pi])
        want, _ = O.run_counts(C.lower_level0(ops, C.SNAKE_LAYOUT), 31, 1, shots=64, noise=onoise, seed=500 + k)
        assert abs(res["autocorr"][pts[k]] - O.compute_z_expectation(want, 1)[0]) < 1e-12, k
    assert np.allclose(res["mean"], res["autocorr"].mean(axis=3))


def test_run_adaptive_equals_oracle_backed_loop(disorder):
    """sweeps.run_adaptive (real-time adaptive g: ctrl-g.py:443-490 feedback, g-opt.py:354-428 grid optimiser) through the
    device simulator == the same control loop on the oracle-backed stand-in simulator: every forward / echo value and the
    whole g history are identical (counts are bit-identical under the shared Philox contract, the control law is
    deterministic).  L = 6 runs the trajectory path (shots < 2^n), L = 4 the exact density-matrix path."""
    from test_dist_cpu import _OracleSim
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    for L, shots, kw in ((6, 64, dict(feedback_gain=0.05, exponential_feedback=True)),
                         (4, 256, dict(feedback_gain=0.05, exponential_feedback=False, g_max=0.95)),
                         (4, 128, dict(use_optimization=True, optimizer="grid", grid_points=4))):
        hs, phis = disorder[20][0][:2, :L], disorder[20][1][:2, :L - 1]
        got = dtcsim.run_adaptive(sim, L, hs, phis, 4, shots=shots, seed_simulator=77, **kw)
        want = dtcsim.run_adaptive(_OracleSim(), L, hs, phis, 4, shots=shots, seed_simulator=77, **kw)
        for key in ("forward", "echo", "g_history"):
            assert np.array_equal(got[key], want[key]), (L, key)
        assert got["circuits"] == want["circuits"]


def test_run_expz_sweep_through_qasm_equals_oracle(disorder):
    """sweeps.run_expz_sweep (dtc_qasm.py:123-160: every (instance, t) circuit handed over as OpenQASM-2 text, L-bit counts,
    per-site <Z>) through the device simulator == the oracle's counts of the same circuits; exact=True agrees with the
    shot estimate within the family-wise bound."""
    L, T = 6, 4
    hs, phis = disorder[20][0][:2, :L], disorder[20][1][:2, :L - 1]
    sim = dtcsim.AerSimulator(noise_model=RC.noise_model(0.05), device="GPU", cuStateVec_enable=True)
    # the untranspiled rx / rzz / rz text carries no u1/u2/u3 gate, so no noise site fires; the oracle gets the same noise
    # model so that both sides select the same method (Aer's automatic rule)
    onoise = O.PauliNoise.depolarizing(0.05)
    res = dtcsim.run_expz_sweep(sim, L, 0.94, hs, phis, T, state="1", shots=512, seed_simulator=40)
    assert res["expz"].shape == (2, L, T - 1)
    k = 0
    for i in range(2):
        for t in range(1, T):
            c = dtcsim.expz_circuit(L, 0.94, hs[i], phis[i], t, "1")
            counts = O.run_counts([o.astuple() for o in c.ops], L, L, shots=512, noise=onoise, seed=40 + k)[0]
            assert np.array_equal(res["expz"][i, :, t - 1], O.compute_z_expectation(counts, L)), (i, t)
            k += 1
    ex = dtcsim.run_expz_sweep(sim, L, 0.94, hs, phis, T, state="1", shots=512, seed_simulator=40, exact=True)
    z = np.abs(ex["expz"] - res["expz"]) / np.sqrt(np.maximum(1 - ex["expz"] ** 2, 1e-3) / 512)
    assert z.max() < family_z(z.size)


def test_thermal_relaxation_through_run_vs_oracle(disorder):
    """Non-Pauli channels of a device-calibrated noise model (thermal relaxation after every u2 / u3, both of Aer's
    constructions: T2 <= T1 mixture incl. reset, T2 > T1 Choi matrix; fast.py:77-78) through run(): exact density matrix
    whatever the shot count, probabilities within 1e-10 of the oracle's Kraus evolution, counts bit-identical; combined with a
    readout error; explicit statevector method refused."""
    from dtcsim import noise as N
    for L, t, echo, (t1, t2), shots in ((4, 3, True, (80.0, 100.0), 1024), (4, 2, False, (100.0, 60.0), 1024),
                                        (6, 2, True, (80.0, 100.0), 16), (7, 1, False, (100.0, 60.0), 64)):
        hs, phis = disorder[20][0][0][:L], disorder[20][1][0][:L - 1]
        circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, t, L // 2, echo))
        nm = dtcsim.NoiseModel()
        nm.add_all_qubit_quantum_error(N.thermal_relaxation_error(t1, t2, 4.0, 0.05), ["u1", "u2", "u3"])
        onoise = O.PauliNoise.thermal_relaxation(t1, t2, 4.0, 0.05)
        sim = dtcsim.AerSimulator(noise_model=nm, device="GPU")
        res = sim.run(circ, shots=shots, seed_simulator=5).result()
        want, info = O.run_counts(RC.ops_of(circ), circ.num_qubits, 1, shots=shots, noise=onoise, seed=5)
        assert res.data()["method"] == "density_matrix" and info["method"] == "density_matrix"
        pr = res.data()["probabilities"]
        assert abs(pr[0] - info["probabilities"][0]) < 1e-10 and abs(sum(pr.values()) - 1) < 1e-12      # keyed by clbit value
        assert res.get_counts() == want, (L, t, echo)
    with pytest.raises(ValueError):
        sim.run(circ, shots=16, method="statevector")
    # amplitude damping on the kick gates only + a readout error on the ancilla
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    circ = RC.transpiled(RC.qc_body("neel", 4, 0.84, hs, phis, 2, 2, True))
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(N.amplitude_damping_error(0.02), ["u3"])
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u2"])
    M = [[0.97, 0.03], [0.08, 0.92]]
    nm.add_all_qubit_readout_error(M)
    ks = [np.array([[1, 0], [0, np.sqrt(0.98)]]), np.array([[0, np.sqrt(0.02)], [0, 0]])]
    onoise = O.PauliNoise({"u2": (0.0125, 0.0125, 0.0125)}, channels={"u3": ks})
    res = dtcsim.AerSimulator(noise_model=nm).run(circ, shots=512, seed_simulator=9).result()
    want, _ = O.run_counts(RC.ops_of(circ), circ.num_qubits, 1, shots=512, noise=onoise, seed=9, readout={0: M})
    assert res.get_counts() == want


def test_readout_errors_counts_vs_oracle(disorder):
    """Classical readout errors (the part of device-calibrated noise, fast.py:77-78, that is not a channel on the state):
    counts bit-identical to the oracle under the shared Philox contract on all three execution paths (density matrix,
    trajectories, ideal multi-clbit), and the reported probabilities are the assignment matrix applied to the exact ones."""
    M = [[0.97, 0.03], [0.08, 0.92]]
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    nm = RC.noise_model(0.05)
    nm.add_all_qubit_readout_error(dtcsim.ReadoutError(M))
    sim = dtcsim.AerSimulator(noise_model=nm, device="GPU", cuStateVec_enable=True)
    onoise = O.PauliNoise.depolarizing(0.05)
    for t, echo, shots in ((2, False, 1024), (3, True, 24)):          # density matrix / trajectories
        circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, t, 2, echo), backend=sim)
        res = sim.run(circ, shots=shots, seed_simulator=77).result()
        want, info = O.run_counts(RC.ops_of(circ), 31, 1, shots=shots, noise=onoise, seed=77, readout={0: M})
        assert res.get_counts() == want, (t, echo, res.get_counts(), want)
        if "probabilities" in info:
            p0, p1 = info["probabilities"]
            assert abs(res.data()["probabilities"][0] - (p0 * M[0][0] + p1 * M[1][0])) < 1e-9
    # readout error only (no gate noise): ideal evolution, flipped records; per-qubit matrices on a measure-all circuit
    ops, n, nc = C.dtc_qasm_gates("1", 6, 0.94, disorder[20][0][0][:6], disorder[20][1][0][:5], 2)
    circ = dtcsim.QuantumCircuit(6, 6)
    for nm_, qs, ps, cs in ops:
        circ._add(nm_, qs, ps, cs)
    ro = dtcsim.NoiseModel()
    ro.add_readout_error(M, [1])
    ro.add_readout_error([[0.9, 0.1], [0.25, 0.75]], [4])
    res = dtcsim.AerSimulator(noise_model=ro).run(circ, shots=2000, seed_simulator=5).result()
    assert res.data()["method"] == "statevector"
    want, _ = O.run_counts(ops, 6, 6, shots=2000, noise=None, seed=5, readout={1: M, 4: [[0.9, 0.1], [0.25, 0.75]]})
    assert res.get_counts() == want
    assert abs(sum(res.data()["probabilities"].values()) - 1) < 1e-12
