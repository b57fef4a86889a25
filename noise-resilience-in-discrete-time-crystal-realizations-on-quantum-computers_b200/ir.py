"""Circuit IR and ingestion for the drop-in ``run()`` boundary.

The reference hands ``backend.run()`` a transpiled qiskit ``QuantumCircuit`` (fast.py:190,211) or,
for dtc_qasm.py, a circuit obtained from an OpenQASM-2 string (dtc_qasm.py:95-107).  qiskit is not
installed in this image, so three front doors lead to the same neutral op list:

* :class:`QuantumCircuit`   -- a small builder with the subset of qiskit's QuantumCircuit API the
  reference scripts call (``h, x, cz, rx, ry, rz, rzz, append, inverse, measure, count_ops`` ...),
  so the scripts' ``qc_qiskit`` bodies (fast.py:124-147) run unchanged against it;
* :func:`from_qiskit`       -- duck-typed reader for a real qiskit circuit (``.data``, ``find_bit``);
* :func:`from_qasm2`        -- reader for the OpenQASM-2 dialect ``qml.to_openqasm`` emits
  (``rx, ry, rz, rzz, x, h, cx, cz, measure``).
"""
import math
import re


class Op:
    __slots__ = ("name", "qubits", "params", "clbits")

    def __init__(self, name, qubits, params=(), clbits=()):
        self.name = name
        self.qubits = tuple(int(q) for q in qubits)
        self.params = tuple(float(p) for p in params)
        self.clbits = tuple(int(c) for c in clbits)

    def astuple(self):
        return (self.name, self.qubits, self.params, self.clbits)

    def __repr__(self):
        return f"Op({self.name}, q={self.qubits}, p={self.params}, c={self.clbits})"


_ONE_Q_NOPARAM = ("h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "sxdg", "id")
_ONE_Q_PARAM = {"rx": 1, "ry": 1, "rz": 1, "u1": 1, "p": 1, "u2": 2, "u3": 3, "u": 3}
_TWO_Q_NOPARAM = ("cx", "cz", "swap")
_TWO_Q_PARAM = {"rzz": 1}
_INVERSE_NAME = {"s": "sdg", "sdg": "s", "t": "tdg", "tdg": "t", "sx": "sxdg", "sxdg": "sx"}


class QuantumCircuit:
    """Minimal qiskit-compatible circuit container (subset used by the reference scripts)."""

    def __init__(self, num_qubits, num_clbits=0, name=None):
        self.num_qubits = int(num_qubits)
        self.num_clbits = int(num_clbits)
        self.name = name or f"circuit-{id(self) & 0xffff}"
        self.ops = []
        self.global_phase = 0.0

    # -- qiskit-like accessors
    @property
    def qubits(self):
        return list(range(self.num_qubits))

    @property
    def clbits(self):
        return list(range(self.num_clbits))

    def count_ops(self):
        c = {}
        for op in self.ops:
            c[op.name] = c.get(op.name, 0) + 1
        return dict(sorted(c.items(), key=lambda kv: -kv[1]))

    def size(self):
        return len(self.ops)

    def copy(self):
        c = QuantumCircuit(self.num_qubits, self.num_clbits, self.name)
        c.ops = list(self.ops)
        c.global_phase = self.global_phase
        return c

    def _add(self, name, qubits, params=(), clbits=()):
        for q in qubits:
            if not 0 <= int(q) < self.num_qubits:
                raise ValueError(f"qubit index {q} out of range for {self.num_qubits}-qubit circuit")
        if len(set(qubits)) != len(qubits):
            raise ValueError(f"duplicate qubits in {name}{tuple(qubits)}")
        self.ops.append(Op(name, qubits, params, clbits))
        return self

    # -- gate methods
    def barrier(self, *qubits):
        return self

    def measure(self, qubit, clbit):
        qs = list(qubit) if hasattr(qubit, "__iter__") else [qubit]
        cs = list(clbit) if hasattr(clbit, "__iter__") else [clbit]
        for q, c in zip(qs, cs):
            if not 0 <= int(c) < self.num_clbits:
                raise ValueError(f"clbit index {c} out of range")
            self._add("measure", (q,), (), (c,))
        return self

    def measure_all(self):
        if self.num_clbits < self.num_qubits:
            self.num_clbits = self.num_qubits
        for q in range(self.num_qubits):
            self._add("measure", (q,), (), (q,))
        return self

    def append(self, sub, qargs=None, cargs=None):
        """Inline another circuit on the listed qubits (fast.py:138 ``circ.append(UF, range(L+1))``)."""
        qargs = list(range(sub.num_qubits)) if qargs is None else [int(q) for q in qargs]
        cargs = list(range(sub.num_clbits)) if cargs is None else [int(c) for c in cargs]
        if len(qargs) != sub.num_qubits:
            raise ValueError("append: qargs length does not match sub-circuit width")
        for op in sub.ops:
            self._add(op.name, [qargs[q] for q in op.qubits], op.params, [cargs[c] for c in op.clbits])
        self.global_phase += sub.global_phase
        return self

    compose = append

    def qasm(self):
        """OpenQASM-2 text of the circuit (the form dtc_qasm.py:95-107 writes to disk and reads back with
        `QuantumCircuit.from_qasm_str`); `from_qasm2(circ.qasm())` returns the same op list.  Parameters are written with
        17 significant digits, so the round trip is exact."""
        lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{self.num_qubits}];"]
        if self.num_clbits:
            lines.append(f"creg c[{self.num_clbits}];")
        for op in self.ops:
            if op.name == "measure":
                lines.append(f"measure q[{op.qubits[0]}] -> c[{op.clbits[0]}];")
                continue
            par = "(" + ",".join(repr(float(x)) for x in op.params) + ")" if op.params else ""
            lines.append(f"{op.name}{par} " + ",".join(f"q[{q}]" for q in op.qubits) + ";")
        return "\n".join(lines) + "\n"

    def inverse(self):
        """Reversed op order with inverted gates (fast.py:141 ``UF_subcircuit.inverse()``)."""
        inv = QuantumCircuit(self.num_qubits, self.num_clbits, self.name + "_dg")
        inv.global_phase = -self.global_phase
        for op in reversed(self.ops):
            nm = op.name
            if nm == "measure":
                raise ValueError("inverse() of a circuit containing measure")
            if nm in ("h", "x", "y", "z", "id", "cx", "cz", "swap"):
                inv._add(nm, op.qubits)
            elif nm in _INVERSE_NAME:
                inv._add(_INVERSE_NAME[nm], op.qubits)
            elif nm in ("rx", "ry", "rz", "u1", "p", "rzz"):
                inv._add(nm, op.qubits, (-op.params[0],))
            elif nm == "u2":
                # u2(phi,lam)^-1 = u3(-pi/2, -lam, -phi)
                inv._add("u3", op.qubits, (-math.pi / 2, -op.params[1], -op.params[0]))
            elif nm in ("u3", "u"):
                inv._add(nm, op.qubits, (-op.params[0], -op.params[2], -op.params[1]))
            else:
                raise ValueError(f"inverse(): unsupported gate {nm}")
        return inv


def _mk1(name):
    def f(self, qubit):
        return self._add(name, (qubit,))
    f.__name__ = name
    return f


def _mk1p(name, npar):
    def f(self, *args):
        return self._add(name, (args[npar],), args[:npar])
    f.__name__ = name
    return f


def _mk2(name):
    def f(self, a, b):
        return self._add(name, (a, b))
    f.__name__ = name
    return f


def _mk2p(name, npar):
    def f(self, *args):
        return self._add(name, (args[npar], args[npar + 1]), args[:npar])
    f.__name__ = name
    return f


for _n in _ONE_Q_NOPARAM:
    setattr(QuantumCircuit, _n, _mk1(_n))
for _n, _k in _ONE_Q_PARAM.items():
    setattr(QuantumCircuit, _n, _mk1p(_n, _k))
for _n in _TWO_Q_NOPARAM:
    setattr(QuantumCircuit, _n, _mk2(_n))
for _n, _k in _TWO_Q_PARAM.items():
    setattr(QuantumCircuit, _n, _mk2p(_n, _k))


# ----------------------------------------------------------------------------- ingestion
def from_qiskit(circ):
    """Duck-typed reader of a qiskit QuantumCircuit (never imports qiskit)."""
    out = QuantumCircuit(circ.num_qubits, circ.num_clbits, getattr(circ, "name", None))
    try:
        out.global_phase = float(getattr(circ, "global_phase", 0.0))
    except TypeError:
        out.global_phase = 0.0
    for inst in circ.data:
        operation = getattr(inst, "operation", None)
        if operation is None:                      # legacy (instruction, qargs, cargs) tuples
            operation, qargs, cargs = inst
        else:
            qargs, cargs = inst.qubits, inst.clbits
        name = operation.name
        if name in ("barrier", "delay"):
            continue
        qs = [circ.find_bit(q).index for q in qargs]
        cs = [circ.find_bit(c).index for c in cargs]
        params = [float(p) for p in operation.params]
        out._add(name, qs, params, cs)
    return out


_QASM_FUNCS = {"pi": math.pi, "sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp,
               "ln": math.log, "sqrt": math.sqrt}


def _eval_param(expr):
    expr = expr.strip().replace("^", "**")
    if not re.fullmatch(r"[0-9eE+\-*/(). a-z_]*", expr):
        raise ValueError(f"qasm: bad parameter expression {expr!r}")
    return float(eval(expr, {"__builtins__": {}}, _QASM_FUNCS))  # noqa: S307 (whitelisted chars)


def from_qasm2(text):
    """Parse the OpenQASM-2 subset ``qml.to_openqasm`` writes (dtc_qasm.py:95-107)."""
    text = re.sub(r"//[^\n]*", "", text)
    stmts = [s.strip() for s in text.replace("\n", " ").split(";") if s.strip()]
    qregs, cregs = {}, {}
    nq = nc = 0
    body = []
    for s in stmts:
        if s.startswith("OPENQASM") or s.startswith("include"):
            continue
        m = re.fullmatch(r"qreg\s+(\w+)\s*\[\s*(\d+)\s*\]", s)
        if m:
            qregs[m.group(1)] = (nq, int(m.group(2)))
            nq += int(m.group(2))
            continue
        m = re.fullmatch(r"creg\s+(\w+)\s*\[\s*(\d+)\s*\]", s)
        if m:
            cregs[m.group(1)] = (nc, int(m.group(2)))
            nc += int(m.group(2))
            continue
        body.append(s)
    circ = QuantumCircuit(nq, nc)

    def bits(tok, regs):
        tok = tok.strip()
        m = re.fullmatch(r"(\w+)\s*\[\s*(\d+)\s*\]", tok)
        if m:
            off, size = regs[m.group(1)]
            i = int(m.group(2))
            if i >= size:
                raise ValueError(f"qasm: index {tok} out of range")
            return [off + i]
        off, size = regs[tok]
        return [off + i for i in range(size)]

    for s in body:
        if s.startswith("barrier"):
            continue
        m = re.fullmatch(r"measure\s+(.+?)\s*->\s*(.+)", s)
        if m:
            for q, c in zip(bits(m.group(1), qregs), bits(m.group(2), cregs)):
                circ.measure(q, c)
            continue
        m = re.fullmatch(r"(\w+)\s*(?:\((.*)\))?\s+(.+)", s)
        if not m:
            raise ValueError(f"qasm: cannot parse statement {s!r}")
        name, pstr, args = m.group(1).lower(), m.group(2), m.group(3)
        params = [_eval_param(p) for p in pstr.split(",")] if pstr else []
        operands = [bits(a, qregs) for a in args.split(",")]
        width = max(len(o) for o in operands)
        for k in range(width):
            qs = [o[k] if len(o) > 1 else o[0] for o in operands]
            if name == "u":
                name = "u3"
            circ._add(name, qs, params)
    return circ


def as_circuit(obj):
    """Accept a native circuit, an OpenQASM-2 string, or a qiskit-like circuit."""
    if isinstance(obj, QuantumCircuit):
        return obj
    if isinstance(obj, str):
        return from_qasm2(obj)
    if hasattr(obj, "data") and hasattr(obj, "find_bit"):
        return from_qiskit(obj)
    raise TypeError(f"cannot interpret {type(obj).__name__} as a circuit")
