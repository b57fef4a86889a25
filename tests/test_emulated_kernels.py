"""The CUDA kernels' per-thread code executed on the CPU (tests/emul) against the oracle.

dtc_hd.cuh / dtc_core.hpp are compiled with g++ and every CTA of k_tile_pass is run thread by thread,
phase by phase.  This validates index maps, swizzles, phase tables, sign masks and the pass scheduler
before any GPU time is spent; the -m gpu tests then check the real kernels through the C ABI.
"""
import numpy as np
import pytest

import dtcsim
import emu
import program_interp as PI
import refcircuits as RC
from dtcsim import compile_circuit
from oracle import oracle as O
from test_planner import _random_circuit


def _check(circ, nm, onoise, seed, n_traj, engine, tol=1e-12, offset=3):
    prog = compile_circuit(circ, nm)
    oc, na, _ = O.compact_ops(RC.ops_of(circ), circ.num_qubits)
    st, fx, fz, ph, npass = emu.run(prog, n_traj=n_traj, traj_offset=offset, seed=seed, engine=engine)
    psi = PI.to_circuit_order(PI.materialize_frame(st, prog.n, fx, fz, ph), prog)
    trajs = np.arange(offset, offset + n_traj)
    ref = O.run_trajectories(oc, na, onoise, seed, trajs) if nm is not None else \
        np.repeat(O.run_statevector(oc, na)[None], n_traj, 0)
    assert np.abs(psi - ref).max() < tol
    return prog, npass


@pytest.mark.parametrize("L,t,echo,pol,layout,state", [
    (11, 2, False, "x", True, "vacuum"),     # n = 12: one tile per state
    (12, 3, True, "x", False, "neel"),       # natural qubit order
    (13, 2, False, "xy", True, "vacuum"),    # two kick layers per period, snake layout
    (13, 2, True, "y", True, "vacuum"),
    (14, 1, True, "yx", True, "neel"),
])
def test_tile_engine_reference_circuits(disorder, L, t, echo, pol, layout, state):
    hs, phis = disorder[20][0][1][:L], disorder[20][1][1][:L - 1]
    circ = RC.transpiled(RC.qc_body(state, L, 0.97, hs, phis, t, L // 2, echo, pol), layout=layout)
    _check(circ, RC.noise_model(0.05), O.PauliNoise.depolarizing(0.05), 77, 3, engine=2)
    _check(circ, None, None, 0, 1, engine=2)


def test_generic_engine_small(disorder):
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, 3, 2, True))
    _check(circ, RC.noise_model(0.05), O.PauliNoise.depolarizing(0.05), 5, 16, engine=1)


@pytest.mark.parametrize("seed", range(4))
def test_tile_engine_random_circuits(seed):
    """General gates: exercises extra / cross / outer two-body terms and odd active sets."""
    rng = np.random.default_rng(100 + seed)
    n = 12 + seed % 2
    circ = _random_circuit(rng, n, 70)
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.3, 1), ["u1", "u2", "u3", "h"])
    onoise = O.PauliNoise.depolarizing(0.3, names=("u1", "u2", "u3", "h"))
    _check(circ, nm, onoise, seed, 2, engine=2, tol=1e-11)


def test_heavy_noise_flips_every_sign(disorder):
    """p = 1 (always a Pauli): every sign mask is exercised."""
    L = 11
    hs, phis = disorder[20][0][2][:L], disorder[20][1][2][:L - 1]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, 2, L // 2, True, "xy"))
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.pauli_error([("X", 0.3), ("Y", 0.3), ("Z", 0.4)]), ["u1", "u2", "u3"])
    onoise = O.PauliNoise({k: (0.3, 0.3, 0.4) for k in ("u1", "u2", "u3")})
    _check(circ, nm, onoise, 11, 3, engine=2)


def test_schedule_one_pass_per_period(disorder):
    """Bulk of the DTC circuit: each fused pass completes one layer (R|B D R'|B then R'|A D' R''|A ...)."""
    L = 20
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, 29, 10, True))
    prog = compile_circuit(circ, RC.noise_model())
    rows, n = emu.schedule(prog)
    assert n <= prog.n_layers + 10, (n, prog.n_layers)
    bulk = rows[(rows[:, 1] >= 0) & (rows[:, 2] >= 0) & (rows[:, 3] >= 0)]
    assert len(bulk) >= 50
    assert (bulk[:, 6] == 0).sum() >= len(bulk) - 8      # chain bonds live in the two tables (no "extra" terms)


def test_sharded_rank_bits(disorder):
    """Top qubits global: diagonal terms on them need no communication (rank bits enter the phases)."""
    rng = np.random.default_rng(5)
    n, n_local = 14, 12
    c = dtcsim.QuantumCircuit(n, 0)
    for layer in range(3):
        for q in range(n_local):
            c.rx(rng.uniform(-3, 3), q)
        for q in range(n - 1):
            c.rzz(rng.uniform(-3, 3), q, q + 1)
        for q in range(n):
            c.rz(rng.uniform(-3, 3), q)
        c.cz(13, 2)
        c.rzz(0.7, 12, 13)
    prog = compile_circuit(c, None, reorder=False)
    ops = RC.ops_of(c)
    for rank in range(4):
        st, fx, fz, ph, _ = emu.run(prog, n_traj=1, n_local=n_local, rank_bits=rank)
        st = PI.materialize_frame(st, n_local, fx, fz, ph)      # frames only touch local qubits here
        ref = O.run_statevector(ops, n, init=rank << n_local)
        assert np.abs(st[0] - ref[rank << n_local:(rank + 1) << n_local]).max() < 1e-12


# ---- streaming engine (k_tile_stream): engine=3 runs eligible passes through its per-thread code
def _stream_equals_tile(prog, n_traj, seed, min_stream=1, **kw):
    s2, fx2, fz2, ph2, _ = emu.run(prog, n_traj=n_traj, seed=seed, traj_offset=5, engine=2, **kw)
    s3, fx3, fz3, ph3, n_stream = emu.run(prog, n_traj=n_traj, seed=seed, traj_offset=5, engine=3, **kw)
    assert n_stream >= min_stream
    assert (fx2 == fx3).all() and (fz2 == fz3).all() and (ph2 == ph3).all()
    assert np.abs(s2 - s3).max() < 1e-13
    return n_stream


@pytest.mark.parametrize("L,t,echo,pol,state", [
    (12, 2, False, "x", "vacuum"),     # factorised register n = 12: one contiguous tile per state (mode A only)
    (13, 2, True, "y", "vacuum"),      # n = 13: [0,12) (mode A) and {0,1} + [3,13) (mode B, 64 B runs)
    (15, 2, True, "xy", "neel"),       # n = 15: mode B with g = 5
    (16, 1, False, "x", "vacuum"),
])
def test_stream_engine_reference_circuits(disorder, L, t, echo, pol, state):
    """Hadamard-test circuits with the ancilla factorised out (what run() executes): every pass streams."""
    hs, phis = disorder[20][0][0][:L], disorder[20][1][0][:L - 1]
    circ = RC.transpiled(RC.qc_body(state, L, 0.97, hs, phis, t, L // 2, echo, pol), layout=True)
    prog = compile_circuit(circ, RC.noise_model(0.05), optimize=True)
    rows, n = emu.schedule(prog)
    assert (rows[:n, 21] > 0).all(), rows[:n, 21]
    if L > 12:
        assert set(rows[:n, 21]) == {1, 2}
    _stream_equals_tile(prog, 3, 91, min_stream=n)
    # full register (ancilla kept): against the oracle, streaming wherever a pass is eligible
    _check(circ, RC.noise_model(0.05), O.PauliNoise.depolarizing(0.05), 91, 2, engine=3)


def test_stream_engine_covers_c2(disorder):
    """Config C2 (L=20, ancilla factorised: n=20): every pass of the schedule is eligible for k_tile_stream."""
    L = 20
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, 29, 10, True))
    prog = compile_circuit(circ, RC.noise_model(), optimize=True)
    rows, n = emu.schedule(prog)
    assert n == prog.n_exec_layers
    assert (rows[:n, 21] > 0).all(), rows[:n, 21]
    assert set(rows[:n, 21]) == {1, 2}


@pytest.mark.parametrize("seed", range(3))
def test_stream_engine_random_circuits(seed):
    """General gates: ineligible passes fall back to the register-fed kernel inside the same run."""
    rng = np.random.default_rng(300 + seed)
    n = 13 + seed
    circ = _random_circuit(rng, n, 60)
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.3, 1), ["u1", "u2", "u3", "h"])
    onoise = O.PauliNoise.depolarizing(0.3, names=("u1", "u2", "u3", "h"))
    _check(circ, nm, onoise, seed, 2, engine=3, tol=1e-11)


def test_stream_engine_heavy_noise(disorder):
    """p = 1 (always a Pauli): every sign mask is exercised on the streaming path."""
    L = 14
    hs, phis = disorder[20][0][2][:L], disorder[20][1][2][:L - 1]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, 2, L // 2, True, "xy"))
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.pauli_error([("X", 0.3), ("Y", 0.3), ("Z", 0.4)]), ["u1", "u2", "u3"])
    prog = compile_circuit(circ, nm, optimize=True)
    _stream_equals_tile(prog, 4, 11, min_stream=3)


def test_stream_engine_sharded_rank_bits():
    rng = np.random.default_rng(6)
    n, n_local = 15, 13
    c = dtcsim.QuantumCircuit(n, 0)
    for layer in range(3):
        for q in range(n_local):
            c.rx(rng.uniform(-3, 3), q)
        for q in range(n - 1):
            c.rzz(rng.uniform(-3, 3), q, q + 1)
        for q in range(n):
            c.rz(rng.uniform(-3, 3), q)
    prog = compile_circuit(c, None, reorder=False)
    ops = RC.ops_of(c)
    for rank in (0, 3):
        st, fx, fz, ph, n_stream = emu.run(prog, n_traj=1, n_local=n_local, rank_bits=rank, engine=3)
        assert n_stream > 0
        st = PI.materialize_frame(st, n_local, fx, fz, ph)
        ref = O.run_statevector(ops, n, init=rank << n_local)
        assert np.abs(st[0] - ref[rank << n_local:(rank + 1) << n_local]).max() < 1e-12


@pytest.mark.parametrize("n,noisy", [(17, True), (18, False)])
def test_stream_engine_mode_c_high_stride_groups(n, noisy):
    """Groups starting at the high-stride bit (lowered to 10 here) get five qubits and tiles of 2 KB runs (mode C)."""
    rng = np.random.default_rng(40 + n)
    c = dtcsim.QuantumCircuit(n, 0)
    for layer in range(3):
        for q in range(n):
            c.rx(rng.uniform(-3, 3), q)
        for q in range(n - 1):
            c.rzz(rng.uniform(-3, 3), q, q + 1)
        for q in range(n):
            c.rz(rng.uniform(-3, 3), q)
    nm = None
    if noisy:
        nm = dtcsim.NoiseModel()
        nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.3, 1), ["rx"])
    prog = compile_circuit(c, nm, reorder=False)
    emu.set_high_stride_bit(10)
    try:
        rows, npass = emu.schedule(prog)
        assert 3 in set(rows[:npass, 21]), rows[:npass, 21]          # mode C passes present
        assert (rows[:npass, 21] > 0).all()
        s2, fx2, fz2, ph2, _ = emu.run(prog, n_traj=2, seed=3, engine=2)
        s3, fx3, fz3, ph3, n_stream = emu.run(prog, n_traj=2, seed=3, engine=3)
    finally:
        emu.set_high_stride_bit(15)
    assert n_stream == npass
    assert np.abs(s2 - s3).max() < 1e-13
    ops = RC.ops_of(c)
    if not noisy:
        psi = PI.materialize_frame(s3, prog.n, fx3, fz3, ph3)
        assert np.abs(psi[0] - O.run_statevector(ops, n)).max() < 1e-11


def test_schedule_large_register_uses_2kb_run_tiles_for_high_groups():
    """n_local = 31 (the sharded L = 34 run): groups [0,10) and [10,20) keep ten qubits (contiguous / 64 B-run tiles), the
    qubits from bit 20 on are swept five at a time in tiles of 32 runs of 2 KB; every pass is eligible for k_tile_stream."""
    rng = np.random.default_rng(7)
    n = 31
    c = dtcsim.QuantumCircuit(n, 0)
    for layer in range(2):
        for q in range(n):
            c.rx(rng.uniform(-3, 3), q)
        for q in range(n - 1):
            c.rzz(rng.uniform(-3, 3), q, q + 1)
        for q in range(n):
            c.rz(rng.uniform(-3, 3), q)
    prog = compile_circuit(c, None, reorder=False)
    rows, npass = emu.schedule(prog)
    rows = rows[:npass]
    assert (rows[:, 21] > 0).all()
    tiles = {tuple(int(x) for x in r[9:21]): int(r[21]) for r in rows}
    assert tiles[tuple(range(12))] == 1
    assert tiles[(0, 1) + tuple(range(10, 20))] == 2
    for g in (20, 25, 26):
        assert tiles[tuple(range(7)) + tuple(range(g, g + 5))] == 3
    assert len(tiles) == 5


# ----------------------------------------------------------------------------------- density matrix (csrc/dtc_dm.cuh)
def _dm_check(circ, nm, tol=1e-13, **opts):
    from dtcsim.backend import flatten_dm_segments
    prog = compile_circuit(circ, nm, want_dm=True, optimize=False)
    rho, info = emu.dm_run(prog.n, flatten_dm_segments(prog.dm_segments), **opts)
    want = PI.run_dm(prog, prog.n)                       # [row, col]
    assert np.abs(rho.T - want).max() < tol
    return prog, info


@pytest.mark.parametrize("L,t,echo,pol,reg,wide13", [
    (5, 2, True, "x", True, True),       # n = 6: one 2^12 tile, three rounds
    (6, 2, False, "xy", True, True),     # n = 7: low group of six + a 2^13 tile with a single qubit (one round, spare bits)
    (7, 1, True, "y", True, True),       # n = 8: 2^13 tile with two qubits
    (7, 1, True, "y", True, False),      # ... 2^12 tiles without a passive bit
    (8, 1, False, "x", True, True),      # n = 9: 2^13 tile with three qubits (two rounds, the last one with spare bits)
    (6, 2, False, "xy", False, True),    # element-per-thread tiles (the emulator's replay of k_dm_tile)
    (4, 3, True, "x", True, True),       # n = 5 (reference config C1): tile smaller than 2^12 -> k_dm_tile
])
def test_dm_passes_reference_circuits(disorder, L, t, echo, pol, reg, wide13):
    """Planner + per-thread code of the density-matrix sweeps == numpy execution of the same segments."""
    hs, phis = disorder[20][0][1][:L], disorder[20][1][1][:L - 1]
    circ = RC.transpiled(RC.qc_body("neel", L, 0.97, hs, phis, t, L // 2, echo, pol))
    _prog, info = _dm_check(circ, RC.noise_model(0.05), reg_passes=reg, wide13=wide13)
    assert (info["reg_passes"] > 0) == (reg and L + 1 >= 6)


@pytest.mark.parametrize("seed", range(3))
def test_dm_passes_random_circuits_asymmetric_channels(seed):
    """General rotations (quarter turns -> swap form, angles beyond pi) and channels with pX != pY != pZ."""
    rng = np.random.default_rng(300 + seed)
    n = 6 + seed
    c = dtcsim.QuantumCircuit(n, 1)
    for _layer in range(3):
        for q in range(n):
            kind = rng.integers(0, 4)
            if kind == 0:
                c.rx(float(rng.uniform(-2 * np.pi, 2 * np.pi)), q)
            elif kind == 1:
                c.rx(float(np.pi * rng.integers(-2, 3)), q)             # exact multiples of pi
            elif kind == 2:
                c.ry(float(rng.uniform(-np.pi, np.pi)), q)
            else:
                c.h(q)
        for q in range(0, n - 1, 2):
            c.rzz(float(rng.uniform(-np.pi, np.pi)), q, q + 1)
        for q in range(1, n - 1, 2):
            c.cz(q, q + 1)
    c.measure(0, 0)
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.pauli_error([("X", 0.07), ("Y", 0.02), ("Z", 0.11), ("I", 0.80)]), ["rx", "ry", "h"])
    _dm_check(c, nm, tol=1e-12)


def test_dm_register_passes_c3_geometry(disorder):
    """BASELINE config C3 (L = 12 ancilla-free chain, rho = 2^24 entries) on the emulator: two register passes per period,
    conflict-free shared-memory accesses, <Z_6> equals the exact light-cone value."""
    from dtcsim.backend import flatten_dm_segments
    hs, phis = disorder[20][0][0][:12], disorder[20][1][0][:11]
    t = 2
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
    prog = compile_circuit(dtcsim.lower_level0(c), RC.noise_model(0.05), want_dm=True, optimize=False)
    rho, info = emu.dm_run(prog.n, flatten_dm_segments(prog.dm_segments))
    assert info["reg_passes"] == 2 * t and info["sweeps"] <= 2 * t + 1 and info["worst_conflict"] == 1
    diag = np.real(np.diagonal(rho))
    assert abs(diag.sum() - 1) < 1e-12
    bit = prog.bit_of[6] if hasattr(prog, "bit_of") else 6
    z = 1.0 - 2.0 * ((np.arange(1 << 12) >> bit) & 1)
    assert abs(float((diag * z).sum()) - O.lightcone_zq(12, 0.97, hs, phis, t, 6, 0.05)) < 1e-10


@pytest.mark.parametrize("seed", range(6))
def test_dm_planner_fuzz_random_segments(seed):
    """Random segment programs straight into the planner: arbitrary qubit subsets per layer (groups that start above qubit 0,
    odd group sizes, single-qubit rounds with spare bits), repeated diagonal segments, channels without a rotation and
    rotations without a channel, quarter / half / full turns.  Register passes and element-per-thread passes against numpy."""
    from types import SimpleNamespace
    from dtcsim.backend import flatten_dm_segments
    rng = np.random.default_rng(900 + seed)
    n = 6 + seed % 4
    segs = []
    for _ in range(6):
        kind = rng.integers(0, 4)
        qs = sorted(rng.choice(n, size=rng.integers(1, n + 1), replace=False).tolist())
        if kind == 0:
            angles = [float(rng.choice([rng.uniform(-7, 7), np.pi / 2, -np.pi / 2, np.pi, 2 * np.pi, 1e-9])) for _ in qs]
            segs.append(("R", list(zip(qs, angles))))
        elif kind == 1:
            segs.append(("N", [(q, tuple(rng.dirichlet([8, 1, 1, 1])[1:])) for q in qs]))
        elif kind == 2:
            segs.append(("R", [(q, float(rng.uniform(-3, 3))) for q in qs]))
            segs.append(("N", [(q, (0.0125, 0.0125, 0.0125)) for q in qs[::2]]))
        else:
            d1 = {int(q): float(rng.uniform(-3, 3)) for q in qs}
            d2 = {(int(a), int(a) + 1): float(rng.uniform(-3, 3)) for a in range(0, n - 1, 2) if rng.random() < 0.7}
            segs.append(("D", d1, d2))
            if rng.random() < 0.3:
                segs.append(("D", {0: 0.3}, {}))
    # a non-trivial start: put every qubit into a superposition first
    segs = [("R", [(q, 0.7 + 0.1 * q) for q in range(n)])] + segs
    want = PI.run_dm(SimpleNamespace(dm_segments=segs), n)
    flat = flatten_dm_segments(segs)
    for reg, wide in ((True, True), (True, False), (False, True)):
        rho, info = emu.dm_run(n, flat, reg_passes=reg, wide13=wide)
        assert np.abs(rho.T - want).max() < 1e-12, (reg, wide)
        assert (info["reg_passes"] > 0) == reg
