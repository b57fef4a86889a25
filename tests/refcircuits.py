"""Reference-style circuit construction through the dtcsim API (test helper).

``qc_body`` is the circuit-building part of the reference's ``qc_qiskit`` (fast.py:125-147,
pol.py:110-155) written against ``dtcsim.QuantumCircuit`` exactly as the scripts write it against
qiskit; ``transpiled`` adds the pass-manager call (fast.py:176-190)."""
import numpy as np

import dtcsim
from dtcsim import QuantumCircuit, generate_preset_pass_manager


def create_UF_subcircuit(L, g, phis, hs, polarization="x"):
    sub = QuantumCircuit(L + 1)
    for i in range(L):
        if polarization == "x":
            sub.rx(np.pi * g, i + 1)
        elif polarization == "y":
            sub.ry(np.pi * g, i + 1)
        elif polarization == "xy":
            sub.rx(np.pi * g / 2, i + 1)
            sub.ry(np.pi * g / 2, i + 1)
        elif polarization == "yx":
            sub.ry(np.pi * g / 2, i + 1)
            sub.rx(np.pi * g / 2, i + 1)
    for i in range(0, L - 1, 2):
        sub.rzz(phis[i], i + 1, i + 2)
    for i in range(1, L - 1, 2):
        sub.rzz(phis[i], i + 1, i + 2)
    for i in range(L):
        sub.rz(hs[i], i + 1)
    return sub


def qc_body(initial_state, L, g, hs, phis, t, qubit, echo=False, polarization="x", g_values=None):
    circ = QuantumCircuit(L + 1, 1)
    if initial_state == "neel":
        for i in range(1, L + 1):
            if i % 2 == 0:
                circ.x(i)
    circ.h(0)
    circ.cz(qubit + 1, 0)
    for step in range(t):
        gg = g if g_values is None else g_values[step]
        circ.append(create_UF_subcircuit(L, gg, phis, hs, polarization), range(L + 1))
    if echo:
        for step in range(t - 1, -1, -1):
            gg = g if g_values is None else g_values[step]
            circ.append(create_UF_subcircuit(L, gg, phis, hs, polarization).inverse(), range(L + 1))
    circ.cz(qubit + 1, 0)
    circ.h(0)
    circ.measure(0, 0)
    return circ


def transpiled(circ, backend=None, layout=True):
    L1 = circ.num_qubits
    init = dtcsim.SNAKE_LAYOUT[:L1] if layout else None
    pm = generate_preset_pass_manager(optimization_level=0, backend=backend, routing_method=None,
                                      initial_layout=init)
    return pm.run(circ)


def noise_model(p=0.05):
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(p, 1), ["u1", "u2", "u3"], warnings=False)
    return nm


def ops_of(circ):
    return [o.astuple() for o in circ.ops]
