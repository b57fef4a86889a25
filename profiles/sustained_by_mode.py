"""Sustained (power-capped) sweep time by tile pattern: n = 20, 1024 trajectories, one program whose rotations sit on the
low group (contiguous tiles) and one on the high group (64 B-run tiles), alternated for `seconds`.  Prints ms/pass of each
and the SM clock seen by nvidia-smi.  Usage: python profiles/sustained_by_mode.py [seconds]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dtcsim  # noqa: E402
from dtcsim import backend, capi  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
n, ntraj = 20, 1024
rng = np.random.default_rng(1)
ctx = backend.DeviceContext(0)
state = ctx.empty(ntraj << n, torch.complex128)
noise = dtcsim.NoiseModel()
noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["rx"])
handles = []
for lo in (0, 10):
    c = dtcsim.QuantumCircuit(n, 0)
    for layer in range(8):
        for q in range(lo, lo + 10):
            c.rx(rng.uniform(-3, 3), q)
        for q in range(n - 1):
            c.rzz(rng.uniform(-3, 3), q, q + 1)
        for q in range(n):
            c.rz(rng.uniform(-3, 3), q)
    prog = dtcsim.compile_circuit(c, dtcsim.as_noise_model(noise), reorder=False)
    h = capi.ProgramHandle(prog, 0)
    h.set_profiling(True)
    handles.append((lo, prog, h))
t0 = time.time()
acc = {0: [0.0, 0], 10: [0.0, 0]}
last = {}
while time.time() - t0 < seconds:
    for lo, prog, h in handles:
        backend.evolve(ctx, prog, ntraj, 0, 1, handle=h, state=state)
        ms, npass = h.pass_time()
        half = sum(h.last_run_flags())
        if time.time() - t0 > seconds / 2:              # second half only: clocks have settled
            acc[lo][0] += ms
            acc[lo][1] += npass - 0.5 * half
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for lo in (0, 10):
    ms, np_ = acc[lo]
    gbs = np_ * 2 * 16 * (1 << n) * ntraj / (ms * 1e-3) / 1e9
    print(f"group [{lo},{lo + 9}]: {ms / np_:.3f} ms per full-traffic pass, {gbs:.0f} GB/s algorithmic (sustained)")
print("nvidia-smi right after:", clk)
