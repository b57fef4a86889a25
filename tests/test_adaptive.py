"""Real-time adaptive control of g (sweeps.run_adaptive / feedback_g / optimize_g; reference: ctrl-g.py:352-398,443-490 and
g-opt.py:354-428,500-556).  CPU: the control law is pinned exactly on the reference's committed g histories; the closed loop
runs on the oracle-backed stand-in simulator; two gloo ranks reproduce one rank."""
import math
import multiprocessing as mp
import os

import numpy as np
import pytest

import dtcsim
from conftest import golden_csv
from test_dist_cpu import _OracleSim, _free_port


@pytest.mark.parametrize("gain,g_max", [("0.01", 1.0), ("0.05", 0.95)])
def test_linear_feedback_reproduces_committed_g_history(gain, g_max):
    """autocorr_data_L4/*realtime_adaptive*gain*.csv: g_{t+1} = clip(g_t + gain (1 - echo_t)) digit for digit."""
    df = golden_csv(f"ref_L4_adaptive_gain{gain}.csv")
    g, echo = list(df["g_history_inst1"]), list(df["echo_adaptive_inst1"])
    assert g[0] == 0.84
    for t in range(len(g) - 1):
        nxt = dtcsim.feedback_g(echo[t], 1.0, g[t], t, float(gain), 0.84, g_max, exponential=False)
        assert abs(nxt - g[t + 1]) < 2e-6, (t, nxt, g[t + 1])


def test_exponential_feedback_branches():
    """calculate_exponential_g_adjustment (ctrl-g.py:352-398) term by term."""
    k, c, t = 0.01, 0.1, 3
    # 0.01 < echo < target: exponential term + 0.1 * log term, scaled by (1 + c t)
    e = 0.4
    want = 0.9 + (k * (1 - e) * math.exp(c * t) + k * math.log(1 / e) * 0.1) * (1 + c * t)
    assert dtcsim.feedback_g(e, 1.0, 0.9, t, k, 0.84, 2.0, True, c) == pytest.approx(want, abs=1e-15)
    # echo above target: no log term
    assert dtcsim.feedback_g(1.2, 1.0, 0.9, t, k, 0.0, 2.0, True, c) == pytest.approx(0.9 + k * (-0.2) * math.exp(c * t) * (1 + c * t), abs=1e-15)
    # echo <= 0.01: strong correction 2 * gain
    assert dtcsim.feedback_g(0.0, 1.0, 0.9, t, k, 0.0, 2.0, True, c) == pytest.approx(0.9 + (k * math.exp(c * t) + 2 * k) * (1 + c * t), abs=1e-15)
    # clipping
    assert dtcsim.feedback_g(0.0, 1.0, 0.99, 10, 0.5, 0.84, 1.0, True, c) == 1.0
    assert dtcsim.feedback_g(5.0, 1.0, 0.85, 0, 0.5, 0.84, 1.0, False) == 0.84


def test_optimize_g_bounded_and_grid():
    calls = []

    def echo_of(cands):
        calls.append(len(cands))
        return [1.0 - 4.0 * (g - 0.9) ** 2 for g in cands]

    g = dtcsim.optimize_g(echo_of, 1.0, 0.84, 1.0, "bounded")
    assert abs(g - 0.9) < 1e-3 and set(calls) == {1}                 # one candidate per evaluation
    calls.clear()
    g = dtcsim.optimize_g(echo_of, 1.0, 0.84, 1.0, "grid", grid_points=9)
    assert calls == [9] and g == pytest.approx(0.9)                  # the whole grid in one call
    # ties: first candidate wins (g-opt.py:419-421 uses a strict <)
    assert dtcsim.optimize_g(lambda c: [0.5] * len(c), 1.0, 0.84, 1.0, "grid", grid_points=5) == 0.84
    with pytest.raises(ValueError):
        dtcsim.optimize_g(echo_of, 1.0, 0.84, 1.0, "newton")


def _disorder4(rows=2):
    import pandas as pd
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    hs = pd.read_csv(os.path.join(g, "hs_L20.csv")).values[:rows, :4]
    phis = pd.read_csv(os.path.join(g, "phis_L20.csv")).values[:rows, :3]
    return hs, phis


def test_closed_loop_is_the_reference_recurrence():
    """Every step's circuits use the g history so far; g follows the measured echo; values = direct evaluation."""
    from oracle import oracle as O
    hs, phis = _disorder4()
    T = 4
    res = dtcsim.run_adaptive(_OracleSim(), 4, hs, phis, T, g_initial=0.84, feedback_gain=0.05, exponential_feedback=False,
                              g_min=0.84, g_max=0.95, shots=256, seed_simulator=11)
    assert res["forward"].shape == res["echo"].shape == res["g_history"].shape == (2, T) and res["circuits"] == 2 * 2 * T
    noise = O.PauliNoise.depolarizing(0.05)
    for i in range(2):
        gh = res["g_history"][i]
        assert gh[0] == 0.84
        for t in range(T):
            if t + 1 < T:
                assert gh[t + 1] == dtcsim.feedback_g(res["echo"][i, t], 1.0, gh[t], t, 0.05, 0.84, 0.95, exponential=False)
            for k, echo in enumerate((False, True)):
                c = dtcsim.autocorr_circuit(4, gh[t], hs[i], phis[i], t + 1, echo=echo, g_values=list(gh[:t + 1]))
                counts = O.run_counts([o.astuple() for o in c.ops], c.num_qubits, c.num_clbits, shots=256, noise=noise,
                                      seed=11 + 1000003 * i + 4099 * t + k)[0]
                z = (counts.get("0", 0) - counts.get("1", 0)) / 256
                assert z == res["echo" if echo else "forward"][i, t]
    assert np.allclose(res["mean_g"], res["g_history"].mean(axis=0))


def test_closed_loop_tracks_committed_history(disorder):
    """Reference run (L = 4, gain 0.05, linear, g in [0.84, 0.95]): our closed loop on the reference's disorder row stays
    within the shot-noise random walk of the committed g history (gain x sigma_echo x sqrt(t) << 0.01)."""
    df = golden_csv("ref_L4_adaptive_gain0.05.csv")
    hs, phis = disorder[4][0][0][:4], disorder[4][1][0][:3]
    T = 8
    res = dtcsim.run_adaptive(_OracleSim(), 4, [hs], [phis], T, g_initial=0.84, feedback_gain=0.05, exponential_feedback=False,
                              g_min=0.84, g_max=0.95, shots=1024, seed_simulator=5)
    assert np.abs(res["g_history"][0] - np.asarray(df["g_history_inst1"][:T])).max() < 0.01


def test_optimization_mode_counts_evaluations():
    hs, phis = _disorder4(1)
    res = dtcsim.run_adaptive(_OracleSim(), 4, hs, phis, 3, g_initial=0.9, use_optimization=True, optimizer="grid", grid_points=4,
                              g_min=0.84, g_max=1.0, shots=128, seed_simulator=3)
    assert res["circuits"] == 2 * 3 + 2 * 4                          # forward + echo per step, one grid after steps 0 and 1
    assert all(g in (0.84, pytest.approx(0.84 + 0.16 / 3), pytest.approx(0.84 + 0.32 / 3), 1.0) for g in res["g_history"][0][1:])


class _TrajSim(_OracleSim):
    """Adds DTCSimulator.sample_trajectories (one Pauli trajectory per shot, shots a..b-1) on top of the oracle."""

    def sample_trajectories(self, circuit, traj_begin, traj_end, seed):
        from oracle import oracle as O
        from oracle import philox
        ops, n, _ = O.compact_ops([o.astuple() for o in circuit.ops], circuit.num_qubits)
        tr = np.arange(traj_begin, traj_end, dtype=np.uint64)
        psi = O.run_trajectories(ops, n, O.PauliNoise.depolarizing(0.05), seed, tr)
        probs = O.outcome_probabilities(np.abs(psi) ** 2, n, O.measured_map(ops), circuit.num_clbits)
        u = philox.uniform(seed, 0, philox.STREAM_MEASURE, tr)
        return np.array([O.sample_outcome(np.cumsum(probs[r]), u[r]) for r in range(len(tr))], dtype=np.int64)


def _adaptive_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    import dtcsim as D
    from test_dist_cpu import _OracleSim as Sim

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    hs, phis = _disorder4(3)
    res = D.run_adaptive(Sim(), 4, hs, phis, 3, feedback_gain=0.05, exponential_feedback=True, shots=64, seed_simulator=21,
                         rank=rank, world=world)
    ez = D.run_expz_sweep(Sim(), 4, 0.94, hs, phis, 3, state="1", shots=64, seed_simulator=8, via_qasm=False,
                          rank=rank, world=world, chunk=2)
    en = D.run_energy_sweep(Sim(), 4, 0.97, hs[:2], phis[:2], [0, 2], precision=1 / 16, seed_simulator=4, rank=rank, world=world)
    from test_adaptive import _TrajSim
    tr = D.run_adaptive(_TrajSim(), 4, hs[:1], phis[:1], 3, feedback_gain=0.05, exponential_feedback=False, shots=48,
                        seed_simulator=2, rank=rank, world=world, parallel="trajectories")
    if rank == 0:
        out.put((res["forward"].tolist(), res["echo"].tolist(), res["g_history"].tolist(), res["circuits"],
                 ez["expz"].tolist(), en["energy_per_site"].tolist(), en["circuits"],
                 tr["echo"].tolist(), tr["g_history"].tolist(), tr["circuits"]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_drivers_two_ranks_equal_one_rank():
    """run_adaptive (instances dealt over two gloo ranks; three instances: ragged), run_expz_sweep and run_energy_sweep
    (points dealt over the ranks) == one rank, value for value."""
    ctx = mp.get_context("spawn")
    results = []
    for world in (1, 2):
        out = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_adaptive_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        results.append(out.get(timeout=300))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    assert results[0] == results[1] and results[0][3] == 3 * 3 * 2
    # the same holds for the site-resolved <Z_i(t)> sweep (6 points dealt over the ranks) and the energy sweep (4 points)
    assert np.array(results[0][4]).shape == (3, 4, 2) and np.array(results[0][5]).shape == (2, 2) and results[0][6] == 8
    # parallel="trajectories": every rank walks the loop, each circuit's shots are split over the ranks (dist.ShardedSampler)
    assert results[0][9] == 6 and np.array(results[0][8]).shape == (1, 3)


# ----------------------------------------------------------------------------------- dtc_qasm.py driver
def test_qasm_export_round_trip(disorder):
    """QuantumCircuit.qasm() -> ir.from_qasm2 returns the same op list, parameters bit for bit (dtc_qasm.py:95-107 hand-over)."""
    hs, phis = disorder[20][0][0][:6], disorder[20][1][0][:5]
    c = dtcsim.expz_circuit(6, 0.94, hs, phis, 3, state="1")
    back = dtcsim.from_qasm2(c.qasm())
    assert back.num_qubits == 6 and back.num_clbits == 6
    assert [o.astuple() for o in back.ops] == [o.astuple() for o in c.ops]
    # the circuit is the oracle's restatement of dtc_qasm.py:70-91, gate for gate
    from oracle import dtc_circuits as C
    ops, n, nc = C.dtc_qasm_gates("1", 6, 0.94, hs, phis, 3)
    assert [o.astuple() for o in c.ops] == [(a, tuple(b), tuple(p), tuple(d)) for a, b, p, d in ops]
    t = dtcsim.autocorr_circuit(4, 0.97, hs, phis, 2, echo=True)          # a transpiled (u2 / u3 / cx / rz) circuit as well
    assert [o.astuple() for o in dtcsim.from_qasm2(t.qasm()).ops] == [o.astuple() for o in t.ops]


def test_run_expz_sweep_layout_and_values(disorder):
    """run_expz_sweep: [instance, site, t-1] from L-bit counts; every entry = the oracle's counts of that point's circuit."""
    from oracle import oracle as O
    L, T = 4, 4
    hs, phis = disorder[20][0][:2, :L], disorder[20][1][:2, :L - 1]
    res = dtcsim.run_expz_sweep(_OracleSim(), L, 0.94, hs, phis, T, state="1", shots=128, seed_simulator=40, via_qasm=False)
    assert res["expz"].shape == (2, L, T - 1) and res["points"] == 2 * (T - 1)
    noise = O.PauliNoise.depolarizing(0.05)
    k = 0
    for i in range(2):
        for t in range(1, T):
            c = dtcsim.expz_circuit(L, 0.94, hs[i], phis[i], t, "1")
            counts = O.run_counts([o.astuple() for o in c.ops], L, L, shots=128, noise=noise, seed=40 + k)[0]
            assert np.array_equal(res["expz"][i, :, t - 1], O.compute_z_expectation(counts, L))
            k += 1
    assert np.allclose(res["mean"], res["expz"].mean(axis=0))


def test_run_shots_sweep_is_one_sweep_per_shot_count(disorder):
    """shots.py: the echo column for each shot count = run_sweep with that many shots and its own seed block."""
    hs, phis = _disorder4()
    tv = [0, 1, 2]
    res = dtcsim.run_shots_sweep(_OracleSim(), 4, 0.84, hs, phis, tv, shot_numbers=(16, 128), seed_simulator=3)
    assert res["shots"] == [16, 128] and res["mean"].shape == (2, 3) and res["autocorr"].shape == (2, 2, 3)
    for k, n in enumerate((16, 128)):
        one = dtcsim.run_sweep(_OracleSim(), 4, [0.84], hs, phis, tv, echoes=(True,), shots=n, seed_simulator=3 + 10007 * k)
        assert np.array_equal(res["autocorr"][k], one["autocorr"][0, 0, 0])
    assert res["periods"] == (16 + 128) * 2 * 2 * sum(tv)
    # multiples of 1 / shots: the estimate really comes from that many samples
    assert np.allclose(res["autocorr"][0] * 16 / 2, np.round(res["autocorr"][0] * 16 / 2))
