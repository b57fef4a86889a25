#!/usr/bin/env python
"""bench_sharded.py -- BASELINE config C5: one kicked-Ising statevector sharded on its top log2(P) qubits.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 \
        bench_sharded.py --L 34 --periods 10

Circuit shape: dtc_qasm.py:70-91 (L qubits, RX(pi g) layer, RZZ even/odd bonds, RZ fields), disorder drawn as
generate_disorder.py:16-18 with default_rng(34), g = 0.97; forward t periods then the inverse t periods
(echo), so <Z_q> = +1 for every qubit is a size-independent correctness check at full size.  Optional
--noise p runs one Pauli trajectory of the forward circuit (frames resolved on the host).  Prints one JSON
line on rank 0 with periods/s, state-sweep and NVLink figures.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def build_circuit(dtcsim, L, g, hs, phis, t, echo):
    c = dtcsim.QuantumCircuit(L, L)
    uf = dtcsim.QuantumCircuit(L)
    for i in range(L):
        uf.rx(np.pi * g, i)
    for i in range(0, L - 1, 2):
        uf.rzz(phis[i], i, i + 1)
    for i in range(1, L - 1, 2):
        uf.rzz(phis[i], i, i + 1)
    for i in range(L):
        uf.rz(hs[i], i)
    for _ in range(t):
        c.append(uf, range(L))
    if echo:
        inv = uf.inverse()
        for _ in range(t):
            c.append(inv, range(L))
    c.measure_all()
    return dtcsim.lower_level0(c)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=34)
    ap.add_argument("--periods", type=int, default=10)
    ap.add_argument("--noise", type=float, default=0.0)
    ap.add_argument("--check-oracle", action="store_true", help="small L only: compare <Z> with the CPU oracle")
    ap.add_argument("--components", action="store_true", help="extra run with a synchronise after every component: seconds each")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import dtcsim
    from dtcsim import sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = args.L
    rng = np.random.default_rng(34)
    hs = rng.random(L) * 2 * np.pi - np.pi
    phis = rng.random(L - 1) * np.pi - 1.5 * np.pi
    g = int(round(np.log2(world)))
    n_local = L - g
    eng = sharded.CudaShardEngine(L, n_local, rank, world, local)

    def allred(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(eng.ctx.device)
        dist.all_reduce(t)
        return t.cpu().numpy()

    out = {}
    # ---- noiseless echo: property check + timing
    circ = build_circuit(dtcsim, L, 0.97, hs, phis, args.periods, True)
    sv = sharded.ShardedStatevector(L, rank, world, eng, all_reduce=allred)
    # warm-up on a one-period circuit (NCCL channels, kernel attributes)
    sharded.ShardedStatevector(L, rank, world, eng, all_reduce=allred).run(
        build_circuit(dtcsim, L, 0.97, hs, phis, 1, True))
    eng.passes = 0
    eng.fast_exchanges = 0
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sv.run(circ)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.ctx.device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    ez = np.array(res["expect_z"])
    periods = 2 * args.periods
    out.update(metric="floquet_periods_per_s", value=periods / dt, unit="periods/s", n_gpus=world,
               config={"workload": f"C5: L={L} statevector sharded on top {g} qubits, {args.periods} periods forward + inverse, complex128",
                       "n_local": n_local, "state_bytes_total": 16 << L},
               seconds=dt, norm=res["norm"], echo_max_dev=float(np.abs(ez - 1).max()),
               segments=sv.stats["segments"], exchanges=sv.stats["exchanges"], exchanges_without_pack=eng.fast_exchanges,
               state_sweeps=eng.passes,
               exchange_bytes_per_rank=sv.stats["exchange_bytes_per_rank"],
               nvlink_gbs_per_rank_lower_bound=sv.stats["exchange_bytes_per_rank"] / dt / 1e9,
               hbm_algorithmic_gbs_per_rank=eng.passes * 2 * (16 << n_local) / dt / 1e9)
    ok = abs(res["norm"] - 1) < 1e-9 and np.abs(ez - 1).max() < 1e-9
    if args.components:
        eng.timing = {}
        sharded.ShardedStatevector(L, rank, world, eng, all_reduce=allred).run(circ)
        out["component_seconds_rank0"] = {k: round(v, 4) for k, v in eng.timing.items()}
        eng.timing = None
    # ---- one noisy forward trajectory
    if args.noise > 0:
        nm = dtcsim.NoiseModel()
        nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(args.noise, 1), ["u1", "u2", "u3"])
        circ_f = build_circuit(dtcsim, L, 0.97, hs, phis, args.periods, False)
        svn = sharded.ShardedStatevector(L, rank, world, eng, all_reduce=allred)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        rn = svn.run(circ_f, nm, seed=1234, trajectory=rank * 0 + 3)
        torch.cuda.synchronize(); dist.barrier()
        out["noisy_trajectory"] = {"seconds": time.perf_counter() - t0, "norm": rn["norm"],
                                   "expect_z_mid": rn["expect_z"][L // 2], "frame_fx": int(rn["frame"][0])}
        ok = ok and abs(rn["norm"] - 1) < 1e-9
        if args.check_oracle and rank == 0:
            from oracle import oracle as O
            oc, na, _ = O.compact_ops([o.astuple() for o in circ_f.ops], L)
            psi = O.run_trajectories(oc, na, O.PauliNoise.depolarizing(args.noise), 1234, [3])[0]
            idx = np.arange(1 << L)
            want = np.array([np.sum(np.abs(psi) ** 2 * (1 - 2 * ((idx >> q) & 1))) for q in range(L)])
            out["oracle_max_dev"] = float(np.abs(want - np.array(rn["expect_z"])).max())
            ok = ok and out["oracle_max_dev"] < 1e-10
    out["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
