"""N > 1 host logic on CPU: world_size-2 gloo process group, trajectory sharding + all-reduce of counts.
The per-range evaluator is the oracle here (the GPU test uses the CUDA simulator through the same code)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import pandas as pd
    import torch.distributed as dist
    from dtcsim import dist as D
    from oracle import dtc_circuits as C
    from oracle import oracle as O
    from oracle import philox

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = os.path.join(root, "tests", "golden")
    hs = pd.read_csv(os.path.join(g, "hs_L20.csv")).values[0][:7]
    phis = pd.read_csv(os.path.join(g, "phis_L20.csv")).values[0][:6]
    ops, _, _ = C.autocorr_gates("vacuum", 7, 0.97, hs, phis, 3, 3, True)
    oc, na, _ = O.compact_ops(C.lower_level0(ops, C.SNAKE_LAYOUT), 31)
    noise = O.PauliNoise.depolarizing(0.05)
    meas = O.measured_map(oc)
    shots, seed = 101, 77

    def evaluate(a, b):
        tr = np.arange(a, b, dtype=np.uint64)
        psi = O.run_trajectories(oc, na, noise, seed, tr)
        probs = O.outcome_probabilities(np.abs(psi) ** 2, na, meas, 1)
        u = philox.uniform(seed, 0, philox.STREAM_MEASURE, tr)
        vals = [O.sample_outcome(np.cumsum(probs[r]), u[r]) for r in range(len(tr))]
        return np.bincount(vals, minlength=2)

    hist = D.sharded_counts(evaluate, shots, 2, rank, world)
    # sweep-level sharding: circuits dealt round-robin, sums all-reduced
    mine = D.deal_units(10, rank, world)
    sums = np.zeros(10)
    for i in mine:
        sums[i] = i * i
    sums = D.all_reduce_sum(sums)
    if rank == 0:
        single = evaluate(0, shots)
        out.put((hist.tolist(), single.tolist(), sums.tolist(), D.counts_dict(hist, 1)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    from dtcsim import dist as D
    for total in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [D.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert sorted(sum((D.deal_units(10, r, 3) for r in range(3)), [])) == list(range(10))


def test_two_rank_gloo_counts_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    hist, single, sums, counts = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert hist == single and sum(hist) == 101           # independent of the number of ranks
    assert sums == [float(i * i) for i in range(10)]
    assert counts == {"0": hist[0], "1": hist[1]}


class _OracleSim:
    """Stands in for DTCSimulator on the CPU: run(list, shots, seed list) -> counts from the oracle (same seeds, same contract)."""

    class _Ctx:
        device = None

    ctx = _Ctx()

    def run(self, circuits, shots=1024, seed_simulator=0, **_kw):
        from oracle import oracle as O
        noise = O.PauliNoise.depolarizing(0.05)
        seeds = list(seed_simulator) if isinstance(seed_simulator, (list, tuple)) else [seed_simulator + i for i in range(len(circuits))]
        outs = [O.run_counts([o.astuple() for o in c.ops], c.num_qubits, c.num_clbits, shots=shots, noise=noise, seed=s)[0]
                for c, s in zip(circuits, seeds)]

        class _R:
            def get_counts(self, i):
                return outs[i]

        class _J:
            def result(self):
                return _R()
        return _J()


def _sweep_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import pandas as pd
    import torch.distributed as dist
    import dtcsim

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    g = os.path.join(root, "tests", "golden")
    hs = pd.read_csv(os.path.join(g, "hs_L20.csv")).values[:2, :4]
    phis = pd.read_csv(os.path.join(g, "phis_L20.csv")).values[:2, :3]
    res = dtcsim.run_sweep(_OracleSim(), 4, [0.84, 0.97], hs, phis, [0, 1, 2], (False, True), ("x", "xy"), shots=32,
                           seed_simulator=9, rank=rank, world=world, chunk=5)
    if rank == 0:
        out.put((res["autocorr"].tolist(), res["points"], res["periods"]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_run_sweep_two_ranks_equal_one_rank():
    """sweeps.run_sweep (the C4 front-end: g x polarisation x echo x instance x t) dealt over 2 gloo ranks == 1 rank:
    per-point seeds are seed + GLOBAL point index, results are all-reduced."""
    ctx = mp.get_context("spawn")
    results = []
    for world in (1, 2):
        out = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        results.append(out.get(timeout=300))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    (a1, n1, per1), (a2, n2, per2) = results
    assert n1 == n2 == 2 * 2 * 2 * 2 * 3 and per1 == per2 == 32 * 2 * 2 * 2 * (3 + 6)
    assert np.array_equal(np.array(a1), np.array(a2))
    assert np.abs(np.array(a1)).max() <= 1.0 and np.abs(np.array(a1)[:, :, :, :, 0] - 0.735).max() < 0.4      # t = 0 rows ~ 0.95^6
