"""Multi-GPU partitioning of the hot path, one process per GPU over torch.distributed.

Two ways the path shards:

* **independent units** -- trajectories of one circuit, and whole circuits (t, echo, disorder instance, g,
  polarisation) of a sweep (fast.py:217-239 loops).  Units are dealt to ranks in contiguous blocks; the
  Philox counter is the *global* trajectory id, so results do not depend on the number of ranks.  The only
  communication is one all-reduce of the count / sum vectors at the end.
* **one large statevector** (n = 34-36) sharded on its top log2(P) qubits -- see sharded.py.

Everything here is host logic over torch tensors; it runs unchanged on CPU tensors with the gloo backend
(tests/test_dist_cpu.py) and on CUDA tensors with NCCL.
"""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous block [a, b) of `total` units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def deal_units(n_units, rank, world):
    """Indices of the units (e.g. circuits of a sweep) owned by `rank`, block-cyclic with block 1."""
    return list(range(rank, n_units, world))


def all_reduce_sum(array, group=None, device=None):
    """Sum a numpy array over ranks (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(array)
    t = torch.as_tensor(np.ascontiguousarray(array))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def sharded_counts(evaluate, shots, n_bins, rank, world, group=None, device=None):
    """Trajectory-sharded counts: `evaluate(a, b)` returns the int64 histogram (length n_bins) of the
    outcomes of trajectories a..b-1; the histograms are summed over ranks (identical on every rank)."""
    a, b = shard_range(shots, rank, world)
    local = np.zeros(n_bins, dtype=np.int64)
    if b > a:
        local += np.asarray(evaluate(a, b), dtype=np.int64)
    return all_reduce_sum(local, group, device)


def counts_dict(hist, n_clbits):
    """Histogram over classical-register values -> Aer-style counts dict (zero-count keys omitted)."""
    return {format(v, f"0{n_clbits}b"): int(c) for v, c in enumerate(hist) if c}


class ShardedSampler:
    """Runs `DTCSimulator`-style trajectory sampling of one circuit across the ranks of a process group.

    sim: a dtcsim.DTCSimulator bound to this rank's GPU.  run() returns the same counts on every rank and
    the same counts a single-GPU run with the same seed returns.
    """

    def __init__(self, sim, rank, world, group=None):
        self.sim, self.rank, self.world, self.group = sim, rank, world, group

    def run_counts(self, circuit, shots=1024, seed_simulator=1234):
        from .ir import as_circuit
        circ = as_circuit(circuit)
        n_bins = 1 << circ.num_clbits

        def evaluate(a, b):
            vals = self.sim.sample_trajectories(circ, a, b, seed_simulator)
            return np.bincount(vals, minlength=n_bins)

        hist = sharded_counts(evaluate, shots, n_bins, self.rank, self.world, self.group, self.sim.ctx.device)
        return counts_dict(hist, circ.num_clbits)
