"""Role timing of k_tile_resident vs k_tile_stream (tuning build: nvcc ... -DDTC_STREAM_TIMING -o csrc/libdtcsim_timing.so).
Usage: DTCSIM_LIB=.../libdtcsim_timing.so python profiles/resident_timing.py [ntraj] [resident MiB]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import dtcsim  # noqa: E402
from dtcsim import backend, capi  # noqa: E402

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 192
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hs, phis = bench.load_disorder(0)
noise = dtcsim.NoiseModel()
noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u1", "u2", "u3"])
circ = bench.qc_circuit(dtcsim, hs, phis, 8, False)
prog = dtcsim.compile_circuit(circ, dtcsim.as_noise_model(noise), optimize=True)
ctx = backend.DeviceContext(0)
lib = capi.load()
capi.set_resident_bytes(mb << 20)
out = (ctypes.c_ulonglong * 16)()
state = ctx.empty(ntraj << prog.n_main, torch.complex128)
for resident in (False, True):
    capi.RESIDENT = resident
    h = capi.ProgramHandle(prog, 0)
    h.set_profiling(True)
    for r in range(3):
        b = backend.evolve(ctx, prog, ntraj, 0, 1 + r, handle=h, state=state, fused_rdm=True)
        ms, n = h.pass_time()
        lib.dtc_debug_stream_timing(out)
    tiles = n * (ntraj << (prog.n_main - 12))
    v = [x / tiles for x in out]
    print(f"resident={b.resident}: {n} sweeps, {ms / n * 1e3:.1f} us/sweep; cycles per tile (summed over CTAs / tiles):")
    print("  TMA driver : wait/poll %.0f, store+wait-read %.0f, load issue %.0f, drain %.0f" % tuple(v[0:4]))
    print("  builder s0 : build1+2 %.0f, wait done %.0f, build3+arrive %.0f, dependency poll %.0f   (per tile of this builder: x3)" % tuple(v[4:8]))
    print("  compute wg0: wait full %.0f, phase1 %.0f, phase2 %.0f, phase3+ %.0f  (per tile of this wg: x2)" % tuple(v[8:12]))
    print("  compute wg1: wait full %.0f, phase1 %.0f, phase2 %.0f, phase3+ %.0f" % tuple(v[12:16]))
    h.close()
