"""Layer compiler: transpiled op stream -> alternating rotation / diagonal layers + noise events.

This is the host half of what qiskit-aer does inside ``backend.run`` (fast.py:211) before it
touches the state: truncate idle qubits (the transpiled circuit is >= 31 qubits wide with only
L+1 active), look up which gates carry noise (fast.py:85-86), and turn the
gate list into device work.  Aer executes gate by gate; we do not.  Every supported gate is
rewritten exactly into two primitives

    ROT(q, theta)         = RX(theta) = exp(-i theta X_q / 2)
    D1(q, a) / D2(i,j,b)  = exp(-i a Z_q / 2), exp(-i b Z_i Z_j / 2)        (diagonal)

using  u3(t,p,l) = e^{i(p+l)/2} RZ(p+pi/2) RX(t) RZ(l-pi/2)  (so the reference's rx -> u3(t,-pi/2,pi/2)
has *no* diagonal part), cz = e^{i pi/4} RZ(pi/2) x RZ(pi/2) x RZZ(-pi/2), cx(c,t) = RY(pi/2)_t CZ RY(-pi/2)_t
and the peephole cx(a,b) rz(p)@b cx(a,b) = RZZ(p) (the transpiler's lowering of rzz, undone).
Primitives are packed into layers  D_0 | R_1 D_1 | R_2 D_2 | ...  such that per-qubit order is
preserved (all D terms commute).  The CUDA engine then streams the state once per R-D-R group.

Pauli noise never touches the state: a sampled Pauli is pushed into a per-trajectory Pauli *frame*
F (psi_true = F psi').  Because every primitive is exp(-i angle P/2) for a Pauli P, conjugation by
the frame only flips the sign of its angle.  The frame kernel therefore walks the EVENT list below
in circuit order and emits one sign bit per event and trajectory; NOISE events update the frame.
A D1 term added after a frame change on its qubit needs its own sign bit, hence two D1 "slots"
per (layer, qubit); if both are taken the qubit is advanced to the next layer.
"""
import math
from collections import defaultdict

import numpy as np

from .ir import QuantumCircuit, as_circuit

PI = math.pi
EV_ROT, EV_D1, EV_D2, EV_NOISE, EV_D2C = 0, 1, 2, 3, 4
VIRTUAL_QUBIT = 63        # table partner of D2C terms: an index bit that is always 0
MAX_D2_PER_LAYER = 64

_U3_OF = {  # name -> (theta, phi, lam, extra global phase) with U = e^{i extra} u3(theta,phi,lam)
    "h": lambda p: (PI / 2, 0.0, PI, 0.0),
    "x": lambda p: (PI, 0.0, PI, 0.0),
    "y": lambda p: (PI, PI / 2, PI / 2, 0.0),
    "u2": lambda p: (PI / 2, p[0], p[1], 0.0),
    "u3": lambda p: (p[0], p[1], p[2], 0.0),
    "u": lambda p: (p[0], p[1], p[2], 0.0),
    "ry": lambda p: (p[0], 0.0, 0.0, 0.0),
}
_DIAG_OF = {  # name -> (rz angle a, global phase) with U = e^{i phase} RZ(a)
    "rz": lambda p: (p[0], 0.0),
    "u1": lambda p: (p[0], p[0] / 2),
    "p": lambda p: (p[0], p[0] / 2),
    "z": lambda p: (PI, PI / 2),
    "s": lambda p: (PI / 2, PI / 4),
    "sdg": lambda p: (-PI / 2, -PI / 4),
    "t": lambda p: (PI / 4, PI / 8),
    "tdg": lambda p: (-PI / 4, -PI / 8),
}
ONE_QUBIT = set(_U3_OF) | set(_DIAG_OF) | {"rx", "sx", "sxdg", "id"}
SUPPORTED = ONE_QUBIT | {"cx", "cz", "rzz", "swap", "measure", "barrier"}


class Program:
    """Compiled circuit: event arrays for the CUDA library + read-out description."""

    def __init__(self):
        self.n = 0                  # all active qubits
        self.n_main = 0             # qubits held in the device register (< n when optimize eliminates some)
        self.n_exec_layers = 1      # layers the device executes; layers beyond belong to prog.small
        self.small = None           # read-out factorisation plan (see compile_circuit)
        self.n_clbits = 0
        self.active = []            # original (wide-circuit) index of each compacted qubit
        self.order = []             # order[bit] = compacted qubit stored at index bit `bit`
        self.bit_of = []            # bit_of[compacted qubit] = index bit
        self.n_layers = 1
        self.ev_type = []
        self.ev_layer = []
        self.ev_q0 = []
        self.ev_q1 = []
        self.ev_slot = []           # D1: slot 0/1; D2: term index within its layer; NOISE: site id
        self.ev_val = []            # ROT: theta; D1: a; D2: b
        self.ev_probs = []          # NOISE: (px, py, pz); zeros otherwise
        self.global_phase = 0.0     # psi_true carries e^{+i global_phase}
        self.measures = []          # (simulated qubit, clbit)
        self.measure_orig = {}      # clbit -> index of the measured qubit in the (wide) input circuit: noise-model key
        self.n_sites = 0
        self.dm_segments = []       # ordered segments for the density-matrix engine
        self.has_channels = False   # non-Pauli channels present: only dm_segments describes the program
        self.rot_layers = 0         # number of non-empty rotation layers (for accounting)

    def arrays(self):
        ne = len(self.ev_type)
        probs = np.zeros((ne, 3), dtype=np.float64)
        for i, p in enumerate(self.ev_probs):
            if p is not None:
                probs[i] = p
        return dict(
            type=np.asarray(self.ev_type, dtype=np.int32), layer=np.asarray(self.ev_layer, dtype=np.int32),
            q0=np.asarray(self.ev_q0, dtype=np.int32), q1=np.asarray(self.ev_q1, dtype=np.int32),
            slot=np.asarray(self.ev_slot, dtype=np.int32), val=np.asarray(self.ev_val, dtype=np.float64),
            probs=probs)

    @property
    def measured_qubits(self):
        return [q for q, _ in self.measures]

    def to_circuit_order(self, psi):
        """Statevector(s) [..., 2^n] in the internal bit order -> compacted circuit-qubit order
        (qubit k of the truncated circuit = bit k), the order Aer's save_statevector would use."""
        n = self.n
        if list(self.order) == list(range(n)):
            return psi
        lead = psi.shape[:-1]
        nl = len(lead)
        v = psi.reshape(lead + (2,) * n)                 # axis nl + n-1-b <-> bit b
        axes = list(range(nl)) + [nl + n - 1 - self.bit_of[n - 1 - k] for k in range(n)]
        return np.ascontiguousarray(np.transpose(v, axes)).reshape(lead + (1 << n,))


def compact(circ):
    """Idle-qubit truncation (what Aer does to the 31-qubit-wide transpiled circuit)."""
    used = sorted({q for op in circ.ops if op.name != "barrier" for q in op.qubits})
    remap = {q: i for i, q in enumerate(used)}
    return remap, used


def choose_bit_order(n, pairs):
    """Internal qubit -> index-bit assignment that keeps interacting qubits close (Cuthill-McKee).

    The simulator is free to store qubit k at any bit of the basis index.  The fused kernel keeps
    two-body phases in two shared-memory tables over *adjacent* tile bits, so a chain should occupy
    consecutive bits (the snake layout of fast.py:177 scatters it).  Pendant qubits hanging off a
    branching node (the Hadamard-test ancilla) go last so they do not split the chain.
    Returns order[bit] = compacted qubit index.
    """
    adj = [set() for _ in range(n)]
    for a, b in pairs:
        if a != b:
            adj[a].add(b)
            adj[b].add(a)
    deg = [len(a) for a in adj]
    pend = [v for v in range(n) if deg[v] == 1 and deg[next(iter(adj[v]))] >= 3]
    core = [v for v in range(n) if v not in pend]
    cdeg = {v: len([w for w in adj[v] if w not in pend]) for v in core}
    seen, order = set(), []
    for start in sorted(core, key=lambda v: (cdeg[v] == 0, 0)):   # keep index order among components
        if start in seen:
            continue
        # component of `start`, begin at its lowest-degree node (lowest index on ties)
        comp, stack = [], [start]
        cs = {start}
        while stack:
            v = stack.pop()
            comp.append(v)
            for w in adj[v]:
                if w not in cs and w not in pend:
                    cs.add(w)
                    stack.append(w)
        root = min(comp, key=lambda v: (cdeg[v], v))
        queue = [root]
        seen.add(root)
        while queue:
            v = queue.pop(0)
            order.append(v)
            for w in sorted((w for w in adj[v] if w not in seen and w not in pend), key=lambda w: (cdeg[w], w)):
                seen.add(w)
                queue.append(w)
    return order + sorted(pend)


def _fuse_cx_rz_cx(ops, noisy):
    """Peephole: cx(a,b) [diag 1q on b] cx(a,b) -> rzz-equivalent; `noisy[i]` must be falsy for all three.

    Only ops touching a or b are neighbours; diagonal 1q gates on the control a in between commute with
    everything involved and stay in place.  Linear time: per-qubit lists of op indices.
    """
    n = len(ops)
    per_q = {}
    where = [None] * n                     # where[i] = {qubit: position of op i in per_q[qubit]}
    for i, op in enumerate(ops):
        w = {}
        for q in op.qubits:
            lst = per_q.setdefault(q, [])
            w[q] = len(lst)
            lst.append(i)
        where[i] = w
    consumed = [False] * n
    out = []
    for i, op in enumerate(ops):
        if consumed[i]:
            continue
        if op.name == "cx" and not noisy[i]:
            a, b = op.qubits
            lb, la = per_q[b], per_q[a]
            kb = where[i][b]
            ok = kb + 2 <= len(lb) - 1
            if ok:
                m, c = lb[kb + 1], lb[kb + 2]
                om, oc = ops[m], ops[c]
                ok = (om.name in _DIAG_OF and om.qubits == (b,) and not noisy[m] and not consumed[m]
                      and oc.name == "cx" and oc.qubits == (a, b) and not noisy[c] and not consumed[c])
                if ok:
                    # everything on the control between the two cx must be a noiseless diagonal 1q gate
                    ka, kc = where[i][a], where[c][a]
                    for k in range(ka + 1, kc):
                        o2 = ops[la[k]]
                        if not (o2.name in _DIAG_OF and o2.qubits == (a,) and not noisy[la[k]]):
                            ok = False
                            break
                if ok:
                    ang, ph = _DIAG_OF[om.name](om.params)
                    out.append(("rzz_fused", (a, b), ang, ph))
                    consumed[m] = consumed[c] = True
                    continue
        out.append(op)
    return out


class _Prims:
    """Lowers gates to the primitive list ("R"|"D"|"N", qubits, value) in circuit order (compacted qubits)."""

    def __init__(self):
        self.prims = []
        self.global_phase = 0.0

    def rot(self, q, theta):
        if theta != 0.0:
            self.prims.append(("R", (q,), theta))

    def d1(self, q, a):
        if a != 0.0:
            self.prims.append(("D", (q,), a))

    def d2(self, i, j, b):
        if b != 0.0:
            self.prims.append(("D", (min(i, j), max(i, j)), b))

    def noise(self, q, probs):
        self.prims.append(("N", (q,), tuple(probs)))

    def channel(self, q, superop):
        self.prims.append(("K", (q,), superop))

    def u3(self, q, theta, phi, lam, extra=0.0):
        self.global_phase += extra + (phi + lam) / 2
        self.d1(q, lam - PI / 2)
        self.rot(q, theta)
        self.d1(q, phi + PI / 2)

    def cz(self, a, b):
        self.global_phase += PI / 4
        self.d1(a, PI / 2)
        self.d1(b, PI / 2)
        self.d2(a, b, -PI / 2)

    def ry(self, q, beta):
        self.d1(q, -PI / 2)
        self.rot(q, beta)
        self.d1(q, PI / 2)

    def cx(self, c, t):
        self.ry(t, -PI / 2)
        self.cz(c, t)
        self.ry(t, PI / 2)


def _is_pauli_rotation(theta):
    """RX(k pi) is a Pauli up to phase: it goes to the frame and leaves psi' untouched."""
    k = round(theta / PI)
    return theta - k * PI == 0.0


class _Layerer:
    """Assigns primitives of one domain (main register or the small read-out simulation) to layers and
    appends the corresponding events to the program."""

    def __init__(self, prog, n):
        self.p = prog
        self.last_r = [0] * n
        self.epoch = [0] * n
        self.d1_slots = {}
        self.d2_terms = {}
        self.d2_count = defaultdict(int)
        self.events = []                      # indices of the events this domain emitted

    def _emit(self, typ, layer, q0, q1, slot, val, probs=None):
        p = self.p
        p.ev_type.append(typ)
        p.ev_layer.append(layer)
        p.ev_q0.append(q0)
        p.ev_q1.append(q1)
        p.ev_slot.append(slot)
        p.ev_val.append(val)
        p.ev_probs.append(probs)
        self.events.append(len(p.ev_type) - 1)
        return len(p.ev_type) - 1

    def rot(self, q, theta):
        if _is_pauli_rotation(theta):
            # RX(k pi) only updates the frame: no state layer, but later D1 terms need a fresh sign slot
            self._emit(EV_ROT, self.last_r[q], q, -1, 0, theta)
            self.epoch[q] += 1
            return
        self.last_r[q] += 1
        self._emit(EV_ROT, self.last_r[q], q, -1, 0, theta)

    def d1(self, q, a):
        while True:
            layer = self.last_r[q]
            slots = self.d1_slots.setdefault((layer, q), [])
            for ep, ev in slots:
                if ep == self.epoch[q]:
                    self.p.ev_val[ev] += a
                    return
            if len(slots) < 2:
                ev = self._emit(EV_D1, layer, q, -1, len(slots), a)
                slots.append((self.epoch[q], ev))
                return
            self.last_r[q] += 1               # both sign slots taken: advance q to the next layer

    def d2(self, i, j, b, classical_partner=False):
        """classical_partner: j is still |0> in psi' -> the term is a one-body phase on i (event D2C)."""
        layer = self.last_r[i] if classical_partner else max(self.last_r[i], self.last_r[j])
        while True:
            key = (layer, i, j, self.epoch[i], self.epoch[j], classical_partner)
            ev = self.d2_terms.get(key)
            if ev is not None:
                self.p.ev_val[ev] += b
                break
            if self.d2_count[layer] < MAX_D2_PER_LAYER:
                ev = self._emit(EV_D2C if classical_partner else EV_D2, layer, i, j, self.d2_count[layer], b)
                self.d2_count[layer] += 1
                self.d2_terms[key] = ev
                break
            layer += 1
        self.last_r[i] = layer
        if not classical_partner:
            self.last_r[j] = layer

    def noise(self, q, probs, site):
        self._emit(EV_NOISE, self.last_r[q], q, -1, site, 0.0, tuple(probs))
        self.epoch[q] += 1


def _analyse_readout(prims, n, measured):
    """Find (E, K, cut): qubits E that never need to enter the big register and the suffix prims[cut:]
    that touches only the small set K (the Hadamard-test ancilla factorises).

    A D2 term whose partner is still classical (never rotated: |0> in psi') is a one-body phase on the
    other qubit.  Qubit a is eliminable if every genuinely two-body term on it lies in a suffix of the
    circuit that touches at most 3 qubits K containing all measured qubits.  Then the state on the other
    qubits is evolved alone, its reduced density matrix on K \\ E is read back, and the few suffix
    operations are applied to a |K|-qubit density matrix per trajectory on the host.
    Returns (E, K_reg, cut, kinds) or None; kinds[idx] = (quantum side, classical partner) for D2 prims.
    """
    classical = [True] * n
    kinds = {}
    for idx, (typ, qs, val) in enumerate(prims):
        if typ == "R":
            if not _is_pauli_rotation(val):
                classical[qs[0]] = False
        elif typ == "D" and len(qs) == 2:
            i, j = qs
            if classical[j]:
                kinds[idx] = (i, j)
            elif classical[i]:
                kinds[idx] = (j, i)
    if not measured or len(set(measured)) > 3:
        return None
    K0 = set(measured)
    partners = set()
    for idx, (typ, qs, val) in enumerate(prims):
        if typ == "D" and len(qs) == 2 and idx not in kinds and (set(qs) & K0):
            partners |= set(qs)
    for K in ([K0 | partners] if len(K0 | partners) <= 3 else []) + [K0]:
        cut = 0
        for idx, (typ, qs, val) in enumerate(prims):
            touched = set(qs) if idx not in kinds else {kinds[idx][0]}
            if touched - K:
                cut = idx + 1
        E = set()
        for a in K:
            ok = True
            for idx, (typ, qs, val) in enumerate(prims[:cut]):
                if a not in qs:
                    continue
                if typ == "D" and len(qs) == 2 and (idx not in kinds or kinds[idx][0] != a):
                    ok = False            # entangling term (or a is the classical side) before the cut
                    break
            if ok:
                E.add(a)
        # a classical partner of an eliminated qubit must not itself be eliminated
        for idx, (side, part) in kinds.items():
            if side in E and part in E:
                E.discard(part)
        K_reg = K - E
        if E and len(K_reg) <= 2 and len(E) < n:
            return sorted(E), sorted(K_reg), cut, kinds
    return None


def _segment_dm(prims):
    """Greedy ASAP packing of primitives into typed segments (R / D / N) preserving per-qubit order."""
    segs = []                                 # [type, payload]
    last_seg = {}
    last_typ = {}
    for typ, qs, val in prims:
        p0 = 0
        for q in qs:
            if q in last_seg:
                share = (typ == "D" and last_typ[q] == "D")
                p0 = max(p0, last_seg[q] if share else last_seg[q] + 1)
        pos = None
        for k in range(p0, len(segs)):
            if segs[k][0] == typ:
                pos = k
                break
        if pos is None:
            segs.append([typ, []])
            pos = len(segs) - 1
        segs[pos][1].append((qs, val))
        for q in qs:
            last_seg[q] = pos
            last_typ[q] = typ
    out = []
    for typ, items in segs:
        if typ == "D":
            d1 = defaultdict(float)
            d2 = defaultdict(float)
            for qs, val in items:
                if len(qs) == 1:
                    d1[qs[0]] += val
                else:
                    d2[qs] += val
            out.append(("D", dict(d1), dict(d2)))
        else:
            out.append((typ, [(qs[0], val) for qs, val in items]))
    return out


def lower_to_prims(circuit, noise_model=None):
    """Stage 1 of the compiler: gates -> primitive list on compacted qubits.

    Returns (prims, global_phase, measured {clbit: compacted qubit}, used (original qubit of each compacted
    one), n_clbits).  prims entries: ("R", (q,), theta) | ("D", (q,), a) | ("D", (i, j), b) | ("N", (q,), (px,py,pz)) |
    ("K", (q,), 4 x 4 superoperator) for a non-Pauli channel (density-matrix programs only).
    """
    circ = as_circuit(circuit)
    for op in circ.ops:
        if op.name not in SUPPORTED:
            raise ValueError(f"unsupported instruction {op.name!r}")
    cmap, used = compact(circ)
    n = len(used)
    if n == 0:
        raise ValueError("circuit has no active qubits")
    if n > 62:
        raise ValueError(f"{n} active qubits exceed the 62-qubit index limit")

    def probs_of(op):
        if noise_model is None or op.name in ("measure", "barrier"):
            return None
        if len(op.qubits) != 1:
            if any(noise_model.lookup(op.name, q) is not None for q in op.qubits):
                raise ValueError(f"noise on multi-qubit gate {op.name!r} is not supported")
            return None
        return noise_model.lookup(op.name, op.qubits[0])

    b = _Prims()
    b.global_phase = float(circ.global_phase)
    ops = [o for o in circ.ops if o.name != "barrier"]
    op_probs = [probs_of(o) for o in ops]
    probs_by_id = {id(o): pr for o, pr in zip(ops, op_probs)}
    ops = _fuse_cx_rz_cx(ops, [pr is not None for pr in op_probs])
    measured = {}
    finished = set()
    for op in ops:
        if isinstance(op, tuple):                      # fused cx-rz-cx
            _, (a, c), ang, ph = op
            qa, qc = cmap[a], cmap[c]
            if qa in finished or qc in finished:
                raise ValueError("gate after measurement (mid-circuit measurement is not supported)")
            b.global_phase += ph
            b.d2(qa, qc, ang)
            continue
        nm = op.name
        qs = [cmap[q] for q in op.qubits]
        if nm == "measure":
            measured[op.clbits[0]] = qs[0]
            finished.add(qs[0])
            continue
        if any(q in finished for q in qs):
            raise ValueError("gate after measurement (mid-circuit measurement is not supported)")
        if nm in _U3_OF:
            th, ph, lam, extra = _U3_OF[nm](op.params)
            b.u3(qs[0], th, ph, lam, extra)
        elif nm in _DIAG_OF:
            ang, ph = _DIAG_OF[nm](op.params)
            b.global_phase += ph
            b.d1(qs[0], ang)
        elif nm == "rx":
            b.rot(qs[0], op.params[0])
        elif nm == "sx":
            b.global_phase += PI / 4
            b.rot(qs[0], PI / 2)
        elif nm == "sxdg":
            b.global_phase -= PI / 4
            b.rot(qs[0], -PI / 2)
        elif nm == "id":
            pass
        elif nm == "rzz":
            b.d2(qs[0], qs[1], op.params[0])
        elif nm == "cz":
            b.cz(qs[0], qs[1])
        elif nm == "cx":
            b.cx(qs[0], qs[1])
        elif nm == "swap":
            b.cx(qs[0], qs[1])
            b.cx(qs[1], qs[0])
            b.cx(qs[0], qs[1])
        pr = probs_by_id[id(op)]
        if pr is not None:
            if isinstance(pr, tuple):
                b.noise(qs[0], pr)
            else:                                       # 4 x 4 superoperator of a non-Pauli channel (noise.ChannelError)
                b.channel(qs[0], pr)
    return b.prims, b.global_phase, measured, used, circ.num_clbits


def compile_circuit(circuit, noise_model=None, want_dm=False, reorder=True, optimize=False):
    """Compile a (transpiled) circuit + noise model into a :class:`Program`.

    optimize=True enables the read-out factorisation (see _analyse_readout): qubits that only couple to
    the rest through a small measured suffix are kept out of the big register (prog.n_main < prog.n)
    and the suffix runs on a tiny density matrix per trajectory (prog.small).  Counts and expectation
    values are unchanged; the full statevector is then not available, so amplitude-level consumers
    compile with optimize=False.

    Raises ValueError for anything the device path cannot execute exactly (unsupported gate,
    noise on a multi-qubit gate, mid-circuit measurement) -- there is no CPU fallback.
    """
    prims, global_phase, measured, used, n_clbits = lower_to_prims(circuit, noise_model)
    n = len(used)
    has_channels = any(typ == "K" for typ, _qs, _v in prims)
    if has_channels and not want_dm:
        raise ValueError("non-Pauli noise channels (thermal relaxation, amplitude damping, Kraus, reset) need the "
                         "density-matrix method: they cannot be sampled as Pauli frames")
    if has_channels and optimize:
        raise ValueError("read-out factorisation is not available with non-Pauli noise channels")

    # ---- stage 2: read-out analysis and internal bit order (eliminated qubits become the top bits)
    plan = _analyse_readout(prims, n, sorted(set(measured.values()))) if optimize else None
    elim = plan[0] if plan else []
    pairs = [qs for typ, qs, _ in prims if typ == "D" and len(qs) == 2]
    order = choose_bit_order(n, pairs) if reorder else list(range(n))
    order = [q for q in order if q not in elim] + [q for q in order if q in elim]
    bit_of = [0] * n
    for bit, cq in enumerate(order):
        bit_of[cq] = bit

    prog = Program()
    prog.n = n
    prog.n_main = n - len(elim)
    prog.active = used
    prog.order = order
    prog.bit_of = bit_of
    prog.n_clbits = n_clbits
    prog.global_phase = global_phase

    # ---- stage 3: primitives -> layered events (main register, then the small read-out domain)
    main = _Layerer(prog, n)
    small = _Layerer(prog, n)
    cut = plan[2] if plan else len(prims)
    kinds = plan[3] if plan else {}
    E = set(elim)
    for idx, (typ, qs, val) in enumerate(prims):
        in_small = bool(plan) and (idx >= cut or bool(set(qs) & E and (idx not in kinds or kinds[idx][0] in E)))
        dom = small if in_small else main
        bq = tuple(bit_of[q] for q in qs)
        if typ == "K":
            continue                                      # density-matrix segments only (prog.has_channels)
        if typ == "R":
            dom.rot(bq[0], val)
        elif typ == "N":
            dom.noise(bq[0], val, prog.n_sites)
            prog.n_sites += 1
        elif len(qs) == 1:
            dom.d1(bq[0], val)
        elif idx in kinds:                                # partner still classical: one-body phase
            side, part = kinds[idx]
            dom.d2(bit_of[side], bit_of[part], val, classical_partner=True)
        else:
            dom.d2(min(bq), max(bq), val)
    prog.n_exec_layers = max(main.last_r) + 1
    n_small_layers = max(small.last_r) + 1 if small.events else 0
    for e in small.events:
        prog.ev_layer[e] += prog.n_exec_layers
    prog.n_layers = prog.n_exec_layers + n_small_layers
    prog.rot_layers = len({prog.ev_layer[e] for e in main.events if prog.ev_type[e] == EV_ROT})
    prog.measures = sorted(((bit_of[q], c) for c, q in measured.items()), key=lambda qc: qc[1])
    prog.measure_orig = {c: used[q] for c, q in measured.items()}
    if plan:
        prog.small = dict(events=sorted(small.events), elim_bits=sorted(bit_of[q] for q in elim),
                          reg_bits=sorted(bit_of[q] for q in plan[1]), first_layer=prog.n_exec_layers,
                          n_layers=n_small_layers)
    prog.has_channels = has_channels                    # the event arrays then omit the channels: rho path only
    if want_dm:
        prog.dm_segments = _segment_dm([(t, tuple(bit_of[q] for q in qs), v) for t, qs, v in prims])
    return prog
