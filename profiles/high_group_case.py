"""Per-group sweep bandwidth of the tile engine on one large state (sharded-run shape): n qubits, rotations on one
group of ten qubits per program.  Usage: python profiles/high_group_case.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dtcsim  # noqa: E402
from dtcsim import backend, capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(1)
ctx = backend.DeviceContext(0)
state = ctx.empty(1 << n, torch.complex128)
for lo in range(0, n, 10):
    qs = list(range(lo, min(lo + 10, n)))
    c = dtcsim.QuantumCircuit(n, 0)
    for layer in range(4):
        for q in qs:
            c.rx(rng.uniform(-3, 3), q)
        for q in range(n - 1):
            c.rzz(rng.uniform(-3, 3), q, q + 1)
        for q in range(n):
            c.rz(rng.uniform(-3, 3), q)
    prog = dtcsim.compile_circuit(c, None, reorder=False)
    h = capi.ProgramHandle(prog, 0)
    h.set_profiling(True)
    for rep in range(2):
        backend.evolve(ctx, prog, 1, 0, 1, handle=h, state=state)
        ms, npass = h.pass_time()
    half = sum(h.last_run_flags())
    gbs = (npass - 0.5 * half) * 2 * 16 * (1 << n) / (ms * 1e-3) / 1e9
    print(f"group [{qs[0]},{qs[-1]}]: {npass} passes ({h.num_stream_passes} streaming, {half} half-traffic) in {ms:.2f} ms -> "
          f"{ms / npass:.2f} ms/pass, {gbs:.0f} GB/s algorithmic")
    h.close()
