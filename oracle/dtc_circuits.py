"""CPU restatement of the reference's circuit construction + level-0 lowering.

TEST INFRASTRUCTURE ONLY.  Produces plain op tuples ``(name, qubits, params, clbits)`` so the
oracle does not depend on the product package.  Each function cites the reference lines it
restates (paths relative to /root/reference, short names as in SURVEY.md):

* ``uf_gates``           <- create_UF_subcircuit  fast.py:111-121, pol.py:110-129, circ-pol.py:110-150
* ``autocorr_gates``     <- qc_qiskit circuit body fast.py:125-147 (time-dependent g: ctrl-g.py:196-241,
                            per-step polarisation: circ-pol.py:164-173, xy-cycle.py:144-156)
* ``dtc_qasm_gates``     <- dtc_qasm.py:70-91 (L-qubit circuit, measure all)
* ``lower_level0``       <- generate_preset_pass_manager(optimization_level=0, ...) fast.py:181-192 as
                            pinned by every committed gate_counts_*aer_simulator*.csv
"""
import math

PI = math.pi


def uf_gates(L, g, phis, hs, polarization="x", time_step=0, circular_frequency=0.0):
    """One Floquet period on system sites (circuit qubits 1..L); high-level gate names."""
    ops = []
    for i in range(L):
        q = i + 1
        if polarization == "x":
            ops.append(("rx", (q,), (PI * g,), ()))
        elif polarization == "y":
            ops.append(("ry", (q,), (PI * g,), ()))
        elif polarization == "xy":
            ops.append(("rx", (q,), (PI * g / 2,), ()))
            ops.append(("ry", (q,), (PI * g / 2,), ()))
        elif polarization == "yx":
            ops.append(("ry", (q,), (PI * g / 2,), ()))
            ops.append(("rx", (q,), (PI * g / 2,), ()))
        elif polarization in ("circular_left", "circular_right"):
            sgn = 1.0 if polarization == "circular_left" else -1.0
            ax = PI * g * math.cos(circular_frequency * time_step) / math.sqrt(2)
            ay = sgn * PI * g * math.sin(circular_frequency * time_step) / math.sqrt(2)
            ops.append(("rx", (q,), (ax,), ()))
            ops.append(("ry", (q,), (ay,), ()))
        elif polarization == "circular_static":
            ops.append(("rx", (q,), (PI * g / math.sqrt(2),), ()))
            ops.append(("ry", (q,), (PI * g / math.sqrt(2),), ()))
        else:
            raise ValueError(polarization)
    for i in range(0, L - 1, 2):
        ops.append(("rzz", (i + 1, i + 2), (float(phis[i]),), ()))
    for i in range(1, L - 1, 2):
        ops.append(("rzz", (i + 1, i + 2), (float(phis[i]),), ()))
    for i in range(L):
        ops.append(("rz", (i + 1,), (float(hs[i]),), ()))
    return ops


def inverse_gates(ops):
    """QuantumCircuit.inverse(): reversed order, negated angles (fast.py:141)."""
    return [(name, qs, tuple(-p for p in params), cs) for (name, qs, params, cs) in reversed(ops)]


def autocorr_gates(initial_state, L, g, hs, phis, t, qubit, echo=False, polarization="x",
                   g_values=None, pol_schedule=None, circular_frequency=0.0):
    """Hadamard-test autocorrelation circuit on L+1 qubits, 1 clbit (fast.py:125-147).

    g_values: optional per-step g list (ctrl-g.py:196-241: step k uses g_values[k], echo undoes
    them in reverse order).  pol_schedule: optional callable step -> polarisation
    (xy-cycle.py:144-156).  Circular polarisations pass time_step=step (circ-pol.py:164-173).
    """
    ops = []
    if initial_state == "neel":
        for i in range(1, L + 1):
            if i % 2 == 0:
                ops.append(("x", (i,), (), ()))
    ops.append(("h", (0,), (), ()))
    ops.append(("cz", (qubit + 1, 0), (), ()))

    def period(step):
        gg = g if g_values is None else g_values[step]
        pol = polarization if pol_schedule is None else pol_schedule(step)
        return uf_gates(L, gg, phis, hs, pol, time_step=step, circular_frequency=circular_frequency)

    for step in range(t):
        ops.extend(period(step))
    if echo:
        for step in range(t - 1, -1, -1):
            ops.extend(inverse_gates(period(step)))
    ops.append(("cz", (qubit + 1, 0), (), ()))
    ops.append(("h", (0,), (), ()))
    ops.append(("measure", (0,), (), (0,)))
    return ops, L + 1, 1


def dtc_qasm_gates(state, L, g, hs, phis, t):
    """dtc_qasm.py:70-91 circuit shape: L qubits, t periods, measure all into c[i]."""
    ops = []
    if state == "1":
        ops.append(("x", (L // 2,), (), ()))
    for _ in range(t):
        for i in range(L):
            ops.append(("rx", (i,), (PI * g,), ()))
        for i in range(0, L - 1, 2):
            ops.append(("rzz", (i, i + 1), (float(phis[i]),), ()))
        for i in range(1, L - 1, 2):
            ops.append(("rzz", (i, i + 1), (float(phis[i]),), ()))
        for i in range(L):
            ops.append(("rz", (i,), (float(hs[i]),), ()))
    for i in range(L):
        ops.append(("measure", (i,), (), (i,)))
    return ops, L, L


def lower_level0(ops, layout=None):
    """Level-0 basis translation to {cx, id, rz, sx, u1, u2, u3} as observed in gate_counts CSVs.

    h -> u2(0,pi); rx(t) -> u3(t,-pi/2,pi/2); ry(t) -> u3(t,0,0); x -> u3(pi,0,pi);
    rzz(p;a,b) -> cx(a,b) rz(p)@b cx(a,b); cz(c,t) -> u2@t cx(c,t) u2@t; rz, measure unchanged.
    layout: optional list mapping circuit qubit k -> physical index (fast.py:176-179).
    """
    m = (lambda q: q) if layout is None else (lambda q: layout[q])
    out = []
    for name, qs, params, cs in ops:
        qs = tuple(m(q) for q in qs)
        if name == "h":
            out.append(("u2", qs, (0.0, PI), ()))
        elif name == "rx":
            out.append(("u3", qs, (params[0], -PI / 2, PI / 2), ()))
        elif name == "ry":
            out.append(("u3", qs, (params[0], 0.0, 0.0), ()))
        elif name == "x":
            out.append(("u3", qs, (PI, 0.0, PI), ()))
        elif name == "rzz":
            a, b = qs
            out.append(("cx", (a, b), (), ()))
            out.append(("rz", (b,), (params[0],), ()))
            out.append(("cx", (a, b), (), ()))
        elif name == "cz":
            c, t = qs
            out.append(("u2", (t,), (0.0, PI), ()))
            out.append(("cx", (c, t), (), ()))
            out.append(("u2", (t,), (0.0, PI), ()))
        elif name in ("rz", "measure", "cx", "u1", "u2", "u3", "id", "barrier"):
            out.append((name, qs, tuple(params), tuple(cs)))
        else:
            raise ValueError(f"lower_level0: no rule for {name}")
    return out


SNAKE_LAYOUT = [15, 30, 17, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 14, 18, 19, 20, 21]  # fast.py:177


def count_ops(ops):
    c = {}
    for name, *_ in ops:
        c[name] = c.get(name, 0) + 1
    return c
