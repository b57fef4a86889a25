"""Per-segment sweep bandwidth of the sharded schedule on ONE GPU (no exchange): n_local = 31 as in L = 34 on 8 GPUs.
T = the top-group segment (R_j|X -> D_j -> R_{j+1}|top group) on the whole 32 GiB shard; S = the rotation-only slice program
on the 4 GiB slices (28 local qubits), whole and group by group.  Usage: python profiles/sharded_pass_case.py [n_local] [g]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dtcsim import backend, capi, sharded  # noqa: E402

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 31
g = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = nl + g
rng = np.random.default_rng(34)
hs = rng.random(n) * 2 * np.pi - np.pi
phis = rng.random(n - 1) * np.pi - 1.5 * np.pi
phys = list(range(n))
ctx = backend.DeviceContext(0)
state = ctx.empty(1 << nl, torch.complex128)
state.zero_()
th = 0.97 * np.pi


def timed(prog, n_local, ptr, reps=3, label=""):
    h = capi.ProgramHandle(prog, 0, capi.ENGINE_AUTO, n_local)
    h.set_profiling(True)
    wsb = h.workspace_bytes(1)
    ws = ctx.empty(wsb, torch.uint8)
    for _ in range(reps):
        h.run(ptr, 1, 0, 0, ws.data_ptr(), wsb, ctx.stream, init_index=capi.INIT_KEEP, rank_bits=0)
        ms, npass = h.pass_time()
    gbs = npass * 2 * 16 * (1 << n_local) / (ms * 1e-3) / 1e9
    print(f"{label:34s} n_local {n_local}: {npass} passes in {ms:8.3f} ms -> {gbs:6.0f} GB/s algorithmic", flush=True)
    h.close()
    return ms


top = max(g, 5)
# T: rotations on the top g bits, full diagonal layer, look-ahead on the top group
T = sharded._SegmentProgram(n, nl)
T.add_layer_rot(1, {q: th for q in range(nl - g, nl)}, phys)
T.add_layer_diag(1, {q: hs[q] for q in range(n)}, {(q, q + 1): phis[q] for q in range(n - 1)}, phys)
T.add_layer_rot(2, {q: th for q in range(nl - top, nl)}, phys)
t_T = timed(T, nl, state.data_ptr(), label="T (top group, diagonal, look-ahead)")
# S: rotations on the bits below the top group, per slice
S = sharded._SegmentProgram(n, nl - g)
S.add_layer_rot(1, {q: th for q in range(0, nl - top)}, phys)
t_S = timed(S, nl - g, state.data_ptr(), label="S (all lower qubits), one slice")
for lo, hi in ((0, 10), (10, 20), (20, 25), (25, nl - top)):
    if hi <= lo:
        continue
    P = sharded._SegmentProgram(n, nl - g)
    P.add_layer_rot(1, {q: th for q in range(lo, hi)}, phys)
    timed(P, nl - g, state.data_ptr(), label=f"  S group [{lo},{hi})")
per = t_T + (1 << g) * t_S
print(f"period without exchange: T {t_T:.2f} ms + {1 << g} slices x {t_S:.2f} ms = {per:.1f} ms")
