"""L2-residency probe for the sweep kernel: the same 9-pass program (L=20 factorised register n=20, forward t=8) on batches
of 2..8 trajectories (32..128 MiB of states: inside / around the 126 MB L2) and on 192 (3 GiB: streams from HBM).
Prints us/pass and algorithmic GB/s per batch size, then board power / SM clock after `seconds` of looping each.
Usage: python profiles/l2_probe.py [seconds per batch size]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import dtcsim  # noqa: E402
from dtcsim import backend, capi  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
hs, phis = bench.load_disorder(0)
noise = dtcsim.NoiseModel()
noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u1", "u2", "u3"])
circ = bench.qc_circuit(dtcsim, hs, phis, 8, False)
prog = dtcsim.compile_circuit(circ, dtcsim.as_noise_model(noise), optimize=True)
ctx = backend.DeviceContext(0)
h = capi.ProgramHandle(prog, 0)
h.set_profiling(True)
state = ctx.empty(192 << prog.n_main, torch.complex128)
Q = "clocks.sm,power.draw"
for ntraj in (2, 3, 4, 5, 6, 8, 16, 192):
    t0 = time.time()
    acc_ms, acc_n, it = 0.0, 0.0, 0
    while time.time() - t0 < seconds:
        for _ in range(20 if ntraj < 100 else 1):
            backend.evolve(ctx, prog, ntraj, 0, 1 + it, handle=h, state=state, fused_rdm=True)
            it += 1
        ms, n = h.pass_time()                           # the last run of the burst
        if time.time() - t0 > seconds / 2:
            acc_ms += ms
            acc_n += n - 0.5 * sum(h.last_run_flags())
    torch.cuda.synchronize()
    clk = subprocess.run(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    gbs = acc_n * 2 * 16 * (1 << prog.n_main) * ntraj / (acc_ms * 1e-3) / 1e9
    print(f"ntraj {ntraj:4d} ({ntraj * 16} MiB): {acc_ms / acc_n * 1e3:8.1f} us per full pass, {gbs:7.0f} GB/s algorithmic; "
          f"{it / (time.time() - t0):7.1f} runs/s; smi: {clk}", flush=True)
