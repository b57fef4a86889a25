"""Level-0 lowering mirror of the reference's transpile step (fast.py:171-192).

The reference calls ``generate_preset_pass_manager(optimization_level=0, backend=backend,
routing_method=None, initial_layout=Layout(snake))`` and hands the result to ``backend.run``.
With qiskit absent this module reproduces what that call does to the reference's circuits, as
pinned by every committed ``gate_counts_*aer_simulator*.csv``:

    h -> u2(0,pi)        rx(t) -> u3(t,-pi/2,pi/2)     ry(t) -> u3(t,0,0)     x -> u3(pi,0,pi)
    rzz(p;a,b) -> cx(a,b) rz(p)@b cx(a,b)               cz(c,t) -> u2@t cx(c,t) u2@t
    rz, u1, u2, u3, cx, id, measure unchanged; no 1q merging, no cancellation.

and the layout widening: circuit qubit k is placed on physical qubit ``initial_layout[k]`` of a
``num_physical``-qubit device.  When qiskit *is* installed the scripts keep using the real pass
manager; :class:`DTCSimulator` exposes the ``target`` it needs.
"""
import math

from .ir import QuantumCircuit

PI = math.pi
SNAKE_LAYOUT = [15, 30, 17, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 14, 18, 19, 20, 21]  # fast.py:177
BASIS_GATES = ("cx", "id", "rz", "sx", "u1", "u2", "u3")


def lower_level0(circ, initial_layout=None, num_physical=None):
    if initial_layout is None:
        layout = list(range(circ.num_qubits))
        width = circ.num_qubits if num_physical is None else num_physical
    else:
        layout = [int(p) for p in initial_layout][:circ.num_qubits]
        if len(layout) < circ.num_qubits:
            raise ValueError("initial_layout shorter than the circuit")
        width = max(max(layout) + 1, num_physical or 0)
    if len(set(layout)) != len(layout):
        raise ValueError("initial_layout maps two circuit qubits to one physical qubit")
    out = QuantumCircuit(width, circ.num_clbits, circ.name)
    out.global_phase = circ.global_phase
    for op in circ.ops:
        qs = [layout[q] for q in op.qubits]
        nm, p = op.name, op.params
        if nm == "h":
            out.u2(0.0, PI, qs[0])
        elif nm == "rx":
            out.u3(p[0], -PI / 2, PI / 2, qs[0])
        elif nm == "ry":
            out.u3(p[0], 0.0, 0.0, qs[0])
        elif nm == "x":
            out.u3(PI, 0.0, PI, qs[0])
        elif nm == "y":
            out.u3(PI, PI / 2, PI / 2, qs[0])
        elif nm == "z":
            out.u1(PI, qs[0])
        elif nm == "s":
            out.u1(PI / 2, qs[0])
        elif nm == "sdg":
            out.u1(-PI / 2, qs[0])
        elif nm == "t":
            out.u1(PI / 4, qs[0])
        elif nm == "tdg":
            out.u1(-PI / 4, qs[0])
        elif nm == "p":
            out.u1(p[0], qs[0])
        elif nm == "u":
            out.u3(p[0], p[1], p[2], qs[0])
        elif nm == "rzz":
            out.cx(qs[0], qs[1])
            out.rz(p[0], qs[1])
            out.cx(qs[0], qs[1])
        elif nm == "cz":
            out.u2(0.0, PI, qs[1])
            out.cx(qs[0], qs[1])
            out.u2(0.0, PI, qs[1])
        elif nm == "swap":
            out.cx(qs[0], qs[1])
            out.cx(qs[1], qs[0])
            out.cx(qs[0], qs[1])
        elif nm in ("rz", "u1", "u2", "u3", "cx", "id", "sx", "measure"):
            out._add(nm, qs, p, op.clbits)
        else:
            raise ValueError(f"lower_level0: no translation rule for gate {nm!r}")
    return out


class _PassManager:
    """Object with the ``.run(circ)`` method the scripts call (fast.py:190)."""

    def __init__(self, initial_layout, num_physical):
        self.initial_layout = initial_layout
        self.num_physical = num_physical

    def run(self, circ):
        return lower_level0(circ, self.initial_layout, self.num_physical)


def generate_preset_pass_manager(optimization_level=0, backend=None, initial_layout=None,
                                 routing_method=None, **_ignored):
    """Same call shape as qiskit's factory for the one configuration the reference uses."""
    if optimization_level != 0:
        raise ValueError("only optimization_level=0 is mirrored (fast.py:171); use qiskit for others")
    nphys = getattr(backend, "num_qubits", None)
    if isinstance(initial_layout, dict):
        initial_layout = [initial_layout[k] for k in sorted(initial_layout)]
    return _PassManager(initial_layout, nphys)
