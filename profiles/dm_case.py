"""Density-matrix sweeps (config C3 shape: n = 12, rho = 2^24 complex128) timed per qubit group.
Usage: [DTCSIM_DM_REG=0|1] [DTCSIM_DM_TILE13=0|1] python profiles/dm_case.py
Prints one JSON line: periods/s of the 20-period C3 program, and microseconds / algorithmic GB/s of a sweep on the
low group (qubits 0-5) and on the high group (qubits 6-11)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dtcsim  # noqa: E402
from dtcsim import backend  # noqa: E402

n, T = 12, 20
rng = np.random.default_rng(5)
hs, phis = rng.uniform(-np.pi, np.pi, n), rng.uniform(-1.5 * np.pi, -0.5 * np.pi, n - 1)
noise = dtcsim.NoiseModel()
noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u1", "u2", "u3"], warnings=False)
ctx = backend.DeviceContext(0)


def circuit(qubits, periods):
    c = dtcsim.QuantumCircuit(n, 1)
    for _ in range(periods):
        for i in qubits:
            c.rx(np.pi * 0.97, i)
        for i in range(0, n - 1, 2):
            c.rzz(phis[i], i, i + 1)
        for i in range(1, n - 1, 2):
            c.rzz(phis[i], i, i + 1)
        for i in range(n):
            c.rz(hs[i], i)
    c.measure(6, 0)
    return dtcsim.compile_circuit(dtcsim.lower_level0(c), dtcsim.as_noise_model(noise), want_dm=True)


def timed(prog, reps):
    stats = {}
    for _ in range(3):
        rho = backend.run_density_matrix(ctx, prog, stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        rho = backend.run_density_matrix(ctx, prog, stats)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, stats["sweeps"], rho


out = {"env": {k: os.environ.get(k) for k in ("DTCSIM_DM_REG", "DTCSIM_DM_TILE13")}}
b_sweep = 2 * 16 * (1 << (2 * n))
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"      # one short run (for ncu)
ms, sweeps, rho = timed(circuit(range(n), 2 if QUICK else T), 1 if QUICK else 10)
d = 1 << n
diag = rho.view(d, d).diagonal().real.cpu().numpy()
out["c3"] = {"periods_per_s": T / (ms * 1e-3), "sweeps": sweeps, "gbs_per_sweep": sweeps * b_sweep / (ms * 1e-3) / 1e9,
             "trace": float(diag.sum()), "z6": float(np.sum(diag * (1.0 - 2.0 * ((np.arange(d) >> 6) & 1))))}
for name, qs in (() if QUICK else (("low_group", range(0, 6)), ("high_group", range(6, 12)))):
    ms, sweeps, _ = timed(circuit(qs, T), 10)
    out[name] = {"sweeps": sweeps, "us_per_sweep": 1e3 * ms / sweeps, "gbs": sweeps * b_sweep / (ms * 1e-3) / 1e9}
print(json.dumps(out), flush=True)
