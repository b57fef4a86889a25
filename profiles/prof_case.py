"""Small, fixed profiling case for ncu: L=20 (n=21) forward t=8 circuit, 96 noisy trajectories (3 GiB of
states, far larger than L2) -> ~20 k_tile_pass launches.  Usage: python profiles/prof_case.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import dtcsim  # noqa: E402
from dtcsim import backend, capi  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ntraj = int(sys.argv[2]) if len(sys.argv) > 2 else 192
hs, phis = bench.load_disorder(0)
noise = dtcsim.NoiseModel()
noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u1", "u2", "u3"])
circ = bench.qc_circuit(dtcsim, hs, phis, 8, False)
prog = dtcsim.compile_circuit(circ, dtcsim.as_noise_model(noise), optimize=True)
ctx = backend.DeviceContext(0)
h = capi.ProgramHandle(prog, 0)
h.set_profiling(True)
state = ctx.empty(ntraj << prog.n_main, torch.complex128)
for r in range(reps):
    b = backend.evolve(ctx, prog, ntraj, 0, 1 + r, handle=h, state=state, fused_rdm=True)
    ms, n = h.pass_time()
    half = sum(h.last_run_flags())          # write-only first pass (generated start), read-only last pass (fused read-out)
    gbs = (n - 0.5 * half) * 2 * 16 * (1 << prog.n_main) * ntraj / (ms * 1e-3) / 1e9
    print(f"rep {r}: {n} passes ({half} half-traffic) in {ms:.3f} ms -> {ms / n * 1e3:.1f} us/pass, {gbs:.0f} GB/s algorithmic")
p = b.outcome_probs().cpu().numpy()
print("mean <Z>", float((p[:, 0] - p[:, 1]).mean()))
