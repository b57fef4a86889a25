#!/usr/bin/env python
"""bench_sharded.py -- BASELINE config C5 alone: the `sharded` sub-record of bench.py for tuning runs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 \
        bench_sharded.py [--periods 10] [--remote-ctas 32] [--no-fuse] [--no-overlap]

Circuit shape: dtc_qasm.py:70-91 (L = 31 + log2 P qubits, RX(pi g) layer, RZZ even/odd bonds, RZ fields), forward t periods
then the inverse t periods (so <Z_q> = +1 for every qubit is a size-independent check at full size), plus an L = 22 noisy
trajectory against the C oracle.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--periods", type=int, default=10)
    ap.add_argument("--remote-ctas", type=int, default=0, help="persistent CTAs of the sweep that stores into the peer (default 64)")
    ap.add_argument("--ce-quarters", type=int, default=-1, help="quarters (0-3) of every slice pushed by the copy engine instead of stored by the sweep")
    ap.add_argument("--no-fuse", action="store_true", help="copy-engine pushes instead of storing the last sweep into the peer")
    ap.add_argument("--no-overlap", action="store_true", help="one all_to_all_single per exchange, no overlap (round-1 behaviour)")
    args = ap.parse_args()
    if args.remote_ctas:
        os.environ["DTCSIM_REMOTE_CTAS"] = str(args.remote_ctas)
    if args.no_fuse:
        os.environ["DTCSIM_FUSE_STORE"] = "0"
    if args.ce_quarters >= 0:
        os.environ["DTCSIM_CE_QUARTERS"] = str(args.ce_quarters)
    import torch
    import torch.distributed as dist
    import bench
    from dtcsim import sharded
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.no_overlap:
        orig = sharded.CudaShardEngine.__init__

        def init(self, *a, **kw):
            kw["overlap"] = False
            kw["transport"] = "nccl"
            orig(self, *a, **kw)
        sharded.CudaShardEngine.__init__ = init
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    ns = argparse.Namespace(sharded_periods=args.periods)
    out = bench.sharded_leg(ns, dist, rank, world, local, peak)
    out["options"] = {"remote_ctas": os.environ.get("DTCSIM_REMOTE_CTAS", "64"), "fuse_store": not args.no_fuse, "ce_quarters": os.environ.get("DTCSIM_CE_QUARTERS", "0"),
                      "overlap": not args.no_overlap}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if out.get("ok") else 1)


if __name__ == "__main__":
    main()
