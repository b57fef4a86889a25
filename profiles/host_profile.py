"""cProfile of the host side of run(): t=29 echo circuit of config C2 with few shots (GPU time negligible)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dtcsim  # noqa: E402

hs, phis = bench.load_disorder(0)
noise = dtcsim.NoiseModel()
noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u1", "u2", "u3"])
circ = bench.qc_circuit(dtcsim, hs, phis, 29, True)
sim = dtcsim.AerSimulator(noise_model=noise)
shots = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sim.run(circ, shots=shots, seed_simulator=1).result()
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    sim.run(circ, shots=shots, seed_simulator=2 + i).result().get_counts()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
