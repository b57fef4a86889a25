"""Small-batch latency of the adaptive loops (ctrl-g.py:464-483, g-opt.py:377-394: 25-40 sequential echo circuits per step,
each ONE circuit of 1024 trajectories whose result decides the next circuit).  Measures one L = 20 echo circuit through run():
cold (first call of the process), warm with a NEW circuit every call (the loops never repeat a circuit: host compile included)
and warm with the same circuit (program cache hit), with the host-side breakdown; under torchrun also
dist.ShardedSampler.run_counts (trajectories split over the ranks + one all-reduce per call).
Usage: python profiles/latency_case.py [t]      or      torchrun --nproc-per-node N profiles/latency_case.py [t]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import dtcsim  # noqa: E402
from dtcsim import capi  # noqa: E402

t = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hs, phis = bench.load_disorder(0)
noise = dtcsim.NoiseModel()
noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u1", "u2", "u3"], warnings=False)
nm = dtcsim.as_noise_model(noise)
sim = dtcsim.AerSimulator(noise_model=noise, device="GPU", cuStateVec_enable=True, cuda_device=local)


def circuit(g):
    return dtcsim.autocorr_circuit(20, g, hs, phis, t, echo=True)


def timed_run(circ, seed):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    counts = sim.run(circ, shots=1024, seed_simulator=seed).result().get_counts()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, counts


out = {"circuit": f"L=20 echo autocorrelation, t={t} ({2 * t} periods), 1024 trajectories, noise p=0.05", "n_gpus": world}
cold_ms, _ = timed_run(circuit(0.97), 1)
out["cold_first_call_ms"] = cold_ms
new = [timed_run(circuit(0.90 + 0.001 * i), 10 + i)[0] for i in range(8)]
same_c = circuit(0.97)
same = [timed_run(same_c, 30 + i)[0] for i in range(8)]
out["warm_new_circuit_ms"] = float(np.median(new))
out["warm_same_circuit_ms"] = float(np.median(same))
# host-side breakdown of a new-circuit call
c = circuit(0.955)
t0 = time.perf_counter()
prog = dtcsim.compile_circuit(c, nm, optimize=True)
t1 = time.perf_counter()
h = capi.ProgramHandle(prog, local)
t2 = time.perf_counter()
from dtcsim import backend  # noqa: E402
ctx = sim.ctx
state = sim._state_buffer(1024 << prog.n_main)
h.set_profiling(True)
torch.cuda.synchronize()
t3 = time.perf_counter()
batch = backend.evolve(ctx, prog, 1024, 0, 5, handle=h, state=state, fused_rdm=True)
probs = batch.outcome_probs()
t4 = time.perf_counter()
torch.cuda.synchronize()
t5 = time.perf_counter()
sweep_ms, n_sweeps = h.pass_time()
out["breakdown_ms"] = {"build_circuit_and_lower": None, "compile_circuit": (t1 - t0) * 1e3, "program_create_upload": (t2 - t1) * 1e3,
                       "enqueue_frames_sweeps_readout": (t4 - t3) * 1e3, "gpu_until_idle": (t5 - t3) * 1e3,
                       "sweep_kernels": sweep_ms, "sweeps": n_sweeps}
t0 = time.perf_counter()
circuit(0.931)
out["breakdown_ms"]["build_circuit_and_lower"] = (time.perf_counter() - t0) * 1e3
h.close()
if dist is not None:
    from dtcsim import dist as D
    sampler = D.ShardedSampler(sim, rank, world)
    sampler.run_counts(circuit(0.97), shots=1024, seed_simulator=3)
    ts = []
    for i in range(8):
        cc = circuit(0.91 + 0.001 * i)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        counts = sampler.run_counts(cc, shots=1024, seed_simulator=50 + i)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    out["sharded_sampler_new_circuit_ms"] = float(np.median(ts))
    out["sharded_sampler_trajectories_per_rank"] = 1024 // world
    ref = sim.run(cc, shots=1024, seed_simulator=57).result().get_counts()
    out["sharded_counts_equal_single_gpu"] = bool(counts == ref)
if rank == 0:
    print(json.dumps(out), flush=True)
if dist is not None:
    dist.destroy_process_group()
