"""Sharded-statevector host logic (sharded.py) on CPU: gloo ranks, numpy local engine, vs the oracle.
Checks frame resolution, layerisation, the one-exchange-per-layer schedule and the qubit permutation."""
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


class NumpyShardEngine:
    """Reference local engine: applies a segment's events gate by gate on the local shard."""

    def __init__(self, n, n_local, rank, world):
        self.n, self.nl, self.rank, self.world = n, n_local, rank, world
        self.psi = np.zeros(1 << n_local, dtype=np.complex128)
        self.exchanges = 0

    def run_segment(self, prog, first):
        if first:
            self.psi[:] = 0
            if self.rank == 0:
                self.psi[0] = 1.0
        self._apply(self.psi, self.nl, self.rank, prog)

    @staticmethod
    def _apply(psi, nl, rank_bits, prog):
        """Events of `prog` on a register of nl local bits whose higher index bits spell rank_bits."""
        ev = prog.arrays()
        idx = np.arange(1 << nl)

        def z(b):
            if b < nl:
                return 1.0 - 2.0 * ((idx >> b) & 1)
            return np.full(1 << nl, 1.0 - 2.0 * ((rank_bits >> (b - nl)) & 1))

        for e in range(len(ev["type"])):
            t, q0, q1, val = int(ev["type"][e]), int(ev["q0"][e]), int(ev["q1"][e]), float(ev["val"][e])
            if t == 0:
                assert q0 < nl, "rotation on a global qubit"
                c, s = np.cos(val / 2), np.sin(val / 2)
                v = psi.reshape(-1, 2, 1 << q0)
                x0, x1 = v[:, 0, :].copy(), v[:, 1, :].copy()
                v[:, 0, :] = c * x0 - 1j * s * x1
                v[:, 1, :] = c * x1 - 1j * s * x0
            elif t == 1:
                psi *= np.exp(-0.5j * val * z(q0))
            else:
                psi *= np.exp(-0.5j * val * z(q0) * z(q1))

    def exchange(self, lq):
        import torch
        import torch.distributed as dist
        g = len(lq)
        idx = np.arange(1 << self.nl)
        d = np.zeros_like(idx)
        for i, b in enumerate(lq):
            d |= ((idx >> b) & 1) << i
        rest = np.zeros_like(idx)
        pos = 0
        for b in range(self.nl):
            if b not in lq:
                rest |= ((idx >> b) & 1) << pos
                pos += 1
        packed = np.empty_like(self.psi)
        packed[(d << (self.nl - g)) | rest] = self.psi
        allp = [torch.zeros(2 << self.nl, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(allp, torch.from_numpy(packed.view(np.float64).copy()))
        chunk = 1 << (self.nl - g)
        recv = np.concatenate([allp[s].numpy().view(np.complex128)[self.rank * chunk:(self.rank + 1) * chunk]
                               for s in range(self.world)])
        self.psi = recv[(d << (self.nl - g)) | rest]
        self.exchanges += 1

    def expect_z_partial(self):
        p = np.abs(self.psi) ** 2
        idx = np.arange(1 << self.nl)
        return np.array([np.sum(p * (1.0 - 2.0 * ((idx >> b) & 1))) for b in range(self.nl)]), float(p.sum())


class NumpySlicedEngine(NumpyShardEngine):
    """Adds the fused form the CUDA engine offers: the program of the lower local qubits is applied slice by slice
    (slice d = what rank d receives, rank bits extended by d), then the top g local bits are exchanged."""

    def __init__(self, *a):
        super().__init__(*a)
        self.sliced = 0

    def exchange_sliced(self, prog, lq, first):
        g = len(lq)
        assert list(lq) == list(range(self.nl - g, self.nl))
        if first:
            self.psi[:] = 0
            if self.rank == 0:
                self.psi[0] = 1.0
        S = 1 << (self.nl - g)
        if prog is not None:
            assert prog.n_main == self.nl - g
            for d in range(1 << g):
                self._apply(self.psi[d * S:(d + 1) * S], self.nl - g, (self.rank << g) | d, prog)
        self.exchange(lq)
        self.sliced += 1


def _worker(rank, world, port, out, sliced=False):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    import dtcsim
    from dtcsim import sharded
    from oracle import dtc_circuits as C
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(34)
    L = 9
    hs = rng.random(L) * 2 * np.pi - np.pi                      # generate_disorder.py:16-18
    phis = rng.random(L - 1) * np.pi - 1.5 * np.pi
    ops, n, nc = C.dtc_qasm_gates("1", L, 0.97, hs, phis, 3)
    low = C.lower_level0(ops)
    circ = dtcsim.QuantumCircuit(L, L)
    for nm, qs, ps, cs in low:
        circ._add(nm, qs, ps, cs)
    noise = dtcsim.NoiseModel()
    noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.2, 1), ["u1", "u2", "u3"])

    def allred(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    res = []
    for traj in (0, 5):
        eng = (NumpySlicedEngine if sliced else NumpyShardEngine)(L, L - int(np.log2(world)), rank, world)
        sv = sharded.ShardedStatevector(L, rank, world, eng, all_reduce=allred)
        r = sv.run(circ, noise, seed=11, trajectory=traj)
        if sliced:
            assert eng.sliced == sv.stats["exchanges"] > 0
        res.append((r["expect_z"], r["norm"], sv.stats["exchanges"], sv.stats["layers"]))
    if rank == 0:
        oc, na, _ = O.compact_ops(low, L)
        want = []
        for traj in (0, 5):
            psi = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(0.2), 11, [traj])[0]
            idx = np.arange(1 << L)
            want.append([float(np.sum(np.abs(psi) ** 2 * (1 - 2 * ((idx >> q) & 1)))) for q in range(L)])
        out.put((res, want))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,sliced", [(2, False), (4, False), (2, True), (4, True)])
def test_sharded_statevector_matches_oracle(world, sliced):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, sliced)) for r in range(world)]
    for p in procs:
        p.start()
    res, want = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (ez, norm, nex, nlay), w in zip(res, want):
        assert abs(norm - 1) < 1e-12
        assert np.abs(np.array(ez) - np.array(w)).max() < 1e-12
        assert nex <= nlay                      # at most one exchange per layer


def test_host_philox_matches_oracle():
    from dtcsim import sharded
    from oracle import philox
    for seed, idx, stream, traj in ((0, 0, 0, 0), (1234, 17, 0, 5), (2 ** 40 + 3, 999, 1, 2 ** 33 + 1)):
        assert sharded.philox_uniform(seed, idx, stream, traj) == float(philox.uniform(seed, idx, stream, traj))
