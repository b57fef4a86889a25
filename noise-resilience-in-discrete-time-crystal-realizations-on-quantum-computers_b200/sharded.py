"""Statevector sharded on its top log2(P) qubits across P ranks (BASELINE config C5).

The reference never simulates more than 21 qubits (its snake layout has 21 entries, fast.py:177); this
module runs the same kicked-Ising circuits (dtc_qasm.py:70-91 shape, or any circuit of the supported gate
set) at n = 34-35, where one state is 256-512 GiB.

Layout: global basis index = (rank << n_local) | local index.  Per layer R_j D_j:
  * diagonal terms need no communication -- the rank bits enter the phases (`rank_bits` of the C ABI);
  * rotations act on local qubits only, so the qubits that are currently global are exchanged with g local
    qubits whose rotation of this layer is already done: pack (k_shard_pack) -> all_to_all_single over
    NVLink -> unpack, and the new qubit->bit permutation is kept (no swap back): ONE exchange per layer.
Noise: the trajectory's Paulis are sampled on the host with the same Philox contract as the device
(`philox_uniform`), and the Pauli frame is tracked on the host while the circuit is cut into per-exchange
segments, so each segment is an ideal program with sign-resolved angles |theta'| <= pi/2.

The local work goes through an engine object (`CudaShardEngine` binds libdtcsim through capi;
tests/test_sharded_cpu.py plugs a numpy engine to check this host logic with gloo on CPU).
"""
import math

import numpy as np

from . import plan

PI = math.pi
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox_uniform(seed, index, stream, traj):
    """Scalar Philox4x32-10 uniform, bit-identical to csrc/dtc_hd.cuh:philox_uniform."""
    c = [index & 0xFFFFFFFF, stream & 0xFFFFFFFF, traj & 0xFFFFFFFF, (traj >> 32) & 0xFFFFFFFF]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return ((c[0] | (c[1] << 32)) >> 11) * (1.0 / 9007199254740992.0)


def resolve_frames(prims, seed, traj):
    """Host frame walk for ONE trajectory: returns sign-resolved ideal primitives (no Pauli rotations, no
    noise) and the final frame (fx, fz, ph).  Same conventions as the device's frame_walk."""
    fx = fz = 0
    ph = 0
    site = 0
    out = []
    for typ, qs, val in prims:
        q = qs[0]
        if typ == "R":
            k = round(val / PI)
            thp = val - k * PI
            if thp != 0.0:
                out.append(("R", qs, -thp if (fz >> q) & 1 else thp))
            if k & 1:
                fx ^= 1 << q
            ph += 3 * k
        elif typ == "D":
            if len(qs) == 1:
                out.append(("D", qs, -val if (fx >> q) & 1 else val))
            else:
                out.append(("D", qs, -val if ((fx >> q) ^ (fx >> qs[1])) & 1 else val))
        else:
            px, py, pz = val
            u = philox_uniform(seed, site, 0, traj)
            site += 1
            fxq = (fx >> q) & 1
            if u < px:
                fx ^= 1 << q
            elif u < px + py:
                ph += 1 + 2 * fxq
                fx ^= 1 << q
                fz ^= 1 << q
            elif u < px + py + pz:
                ph += 2 * fxq
                fz ^= 1 << q
    return out, fx, fz, ph % 4


def layerize(prims, n):
    """Ideal primitives -> layers [(rot {q: theta}, d1 {q: a}, d2 {(i,j): b})], per-qubit order preserved."""
    last_r = [0] * n
    layers = [({}, {}, {})]

    def layer(j):
        while len(layers) <= j:
            layers.append(({}, {}, {}))
        return layers[j]

    for typ, qs, val in prims:
        if typ == "R":
            last_r[qs[0]] += 1
            layer(last_r[qs[0]])[0][qs[0]] = val
        elif len(qs) == 1:
            d = layer(last_r[qs[0]])[1]
            d[qs[0]] = d.get(qs[0], 0.0) + val
        else:
            j = max(last_r[qs[0]], last_r[qs[1]])
            d = layer(j)[2]
            d[qs] = d.get(qs, 0.0) + val
            last_r[qs[0]] = last_r[qs[1]] = j
    return layers


class _SegmentProgram:
    """Duck-typed stand-in for plan.Program holding one segment's events on *physical* bit indices."""

    def __init__(self, n, n_local):
        self.n, self.n_main = n, n_local
        self.n_layers = self.n_exec_layers = 1
        self.global_phase = 0.0
        self.ev = {k: [] for k in ("type", "layer", "q0", "q1", "slot", "val")}
        self._terms = {}

    def add_layer_rot(self, layer, rot, phys):
        for q, th in rot.items():
            self._emit(plan.EV_ROT, layer, phys[q], -1, 0, th)
        self.n_layers = self.n_exec_layers = max(self.n_layers, layer + 1)

    def add_layer_diag(self, layer, d1, d2, phys):
        for q, a in d1.items():
            if a != 0.0:
                self._emit(plan.EV_D1, layer, phys[q], -1, 0, a)
        for (i, j), b in d2.items():
            if b != 0.0:
                k = self._terms.get(layer, 0)
                if k >= plan.MAX_D2_PER_LAYER:
                    raise ValueError("more than 64 two-body terms in one layer of a sharded run")
                self._terms[layer] = k + 1
                self._emit(plan.EV_D2, layer, phys[i], phys[j], k, b)
        self.n_layers = self.n_exec_layers = max(self.n_layers, layer + 1)

    def _emit(self, typ, layer, q0, q1, slot, val):
        for k, v in zip(("type", "layer", "q0", "q1", "slot", "val"), (typ, layer, q0, q1, slot, val)):
            self.ev[k].append(v)

    def arrays(self):
        ne = len(self.ev["type"])
        return dict(type=np.asarray(self.ev["type"], dtype=np.int32), layer=np.asarray(self.ev["layer"], dtype=np.int32),
                    q0=np.asarray(self.ev["q0"], dtype=np.int32), q1=np.asarray(self.ev["q1"], dtype=np.int32),
                    slot=np.asarray(self.ev["slot"], dtype=np.int32), val=np.asarray(self.ev["val"], dtype=np.float64),
                    probs=np.zeros((ne, 3), dtype=np.float64))


class ShardedStatevector:
    """Evolves one n-qubit state over `world` = 2^g ranks.  engine: object with
    run_segment(prog, first), exchange(lq), expect_z_partial() -> (sum |psi|^2 z_b for local bits b, local norm)."""

    def __init__(self, n, rank, world, engine, all_reduce=None):
        g = int(round(math.log2(world)))
        if 1 << g != world:
            raise ValueError("world size must be a power of two")
        self.n, self.g, self.rank, self.world = n, g, rank, world
        self.n_local = n - g
        if self.n_local < g:
            raise ValueError("too few local qubits for this many ranks")
        self.engine = engine
        self.all_reduce = all_reduce or (lambda a: a)
        self.stats = dict(segments=0, exchanges=0, exchange_bytes_per_rank=0, layers=0)

    def run(self, circuit, noise_model=None, seed=0, trajectory=0):
        """Returns {'expect_z': [<Z_q>] per compacted circuit qubit, 'norm': float, 'frame': (fx, fz, ph)}."""
        prims, _gp, _meas, used, _ncl = plan.lower_to_prims(circuit, noise_model)
        n = len(used)
        if n != self.n:
            raise ValueError(f"circuit has {n} active qubits, the sharded register has {self.n}")
        ideal, fx, fz, ph = resolve_frames(prims, seed, trajectory)
        layers = layerize(ideal, n)
        self.stats["layers"] = len(layers)
        nl, g = self.n_local, self.g
        phys = list(range(n))                      # logical qubit -> physical bit; bits >= n_local are global
        first = True
        applied = set()
        j = 0
        M = len(layers)
        while j < M:
            rot, d1, d2 = layers[j]
            todo = {q: th for q, th in rot.items() if q not in applied}
            glob = [q for q in todo if phys[q] >= nl]
            if glob:
                loc = {q: th for q, th in todo.items() if phys[q] < nl}
                if loc:
                    prog = _SegmentProgram(n, nl)
                    prog.add_layer_rot(1, loc, phys)
                    self.engine.run_segment(prog, first)
                    first = False
                    self.stats["segments"] += 1
                    applied |= set(loc)
                # swap the g global bits with g local qubits whose layer-j rotation is done (or that have none)
                cand = [q for q in range(n) if phys[q] < nl and (q in applied or q not in rot)]
                if len(cand) < g:
                    raise ValueError("not enough finished local qubits to exchange with the global ones")
                cand.sort(key=lambda q: -phys[q])
                out_q = cand[:g]
                lq = sorted(phys[q] for q in out_q)
                self.engine.exchange(lq)
                self.stats["exchanges"] += 1
                self.stats["exchange_bytes_per_rank"] += (16 << nl) * (self.world - 1) // self.world
                inv = {phys[q]: q for q in range(n)}
                for i, b in enumerate(lq):
                    ql, qg = inv[b], inv[nl + i]
                    phys[ql], phys[qg] = nl + i, b
                continue
            prog = _SegmentProgram(n, nl)
            prog.add_layer_rot(1, todo, phys)
            prog.add_layer_diag(1, d1, d2, phys)
            nxt = {}
            if j + 1 < M:
                nxt = {q: th for q, th in layers[j + 1][0].items() if phys[q] < nl}
                prog.add_layer_rot(2, nxt, phys)
            self.engine.run_segment(prog, first)
            first = False
            self.stats["segments"] += 1
            applied = set(nxt)
            j += 1
        zsum, norm = self.engine.expect_z_partial()           # local bits; global bits follow from the rank
        tot = np.zeros(n + 1)
        for q in range(n):
            b = phys[q]
            tot[q] = zsum[b] if b < nl else norm * (1.0 - 2.0 * ((self.rank >> (b - nl)) & 1))
        tot[n] = norm
        tot = self.all_reduce(tot)
        ez = [(-tot[q] if (fx >> q) & 1 else tot[q]) / tot[n] for q in range(n)]
        return {"expect_z": ez, "norm": float(tot[n]), "frame": (fx, fz, ph), "phys": list(phys)}


class CudaShardEngine:
    """Local shard on one GPU: libdtcsim for the fused passes, k_shard_pack + all_to_all_single for exchanges."""

    def __init__(self, n, n_local, rank, world, device_index, group=None):
        import torch
        from . import backend, capi
        self.torch, self.capi = torch, capi
        self.n, self.n_local, self.rank, self.world, self.group = n, n_local, rank, world, group
        self.ctx = backend.DeviceContext(device_index)
        self.a = self.ctx.empty(1 << n_local, torch.complex128)
        self.b = self.ctx.empty(1 << n_local, torch.complex128)
        self.passes = 0
        self.fast_exchanges = 0
        self.timing = None          # set to {} to collect wall-clock seconds per component (synchronises after each)

    def _timed(self, key, t0):
        if self.timing is not None:
            import time
            self.torch.cuda.synchronize(self.ctx.index)
            self.timing[key] = self.timing.get(key, 0.0) + time.perf_counter() - t0

    def run_segment(self, prog, first):
        import time
        capi = self.capi
        t0 = time.perf_counter()
        h = capi.ProgramHandle(prog, self.ctx.index, capi.ENGINE_AUTO, self.n_local)
        self._timed("segment_setup", t0)
        t0 = time.perf_counter()
        wsb = h.workspace_bytes(1)
        ws = self.ctx.empty(wsb, self.torch.uint8)
        init = capi.INIT_KEEP if not first else (0 if self.rank == 0 else capi.INIT_ZERO)
        h.run(self.a.data_ptr(), 1, 0, 0, ws.data_ptr(), wsb, self.ctx.stream, init_index=init, rank_bits=self.rank)
        self.passes += h.num_passes
        self.torch.cuda.current_stream(self.ctx.index).synchronize()     # ws / handle are freed on return
        self._timed("segment_sweeps", t0)
        t0 = time.perf_counter()
        h.close()
        self._timed("segment_teardown", t0)

    def exchange(self, lq):
        import torch.distributed as dist
        capi, lib = self.capi, self.capi.load()
        g = len(lq)
        f64 = self.torch.float64
        if list(lq) == list(range(self.n_local - g, self.n_local)):
            # the outgoing qubits are the top local bits: chunk d of the state IS what rank d receives and the incoming
            # chunks land where they belong, so pack and unpack are identities -- one all-to-all, no extra sweeps
            import time
            t0 = time.perf_counter()
            dist.all_to_all_single(self.b.view(f64), self.a.view(f64), group=self.group)
            self.a, self.b = self.b, self.a
            self.fast_exchanges += 1
            self._timed("all_to_all", t0)
            return
        _, lp = capi.i32(lq)
        capi.check(lib.dtc_shard_pack(self.a.data_ptr(), self.b.data_ptr(), self.n_local, g, lp, self.ctx.stream))
        dist.all_to_all_single(self.a.view(f64), self.b.view(f64), group=self.group)
        capi.check(lib.dtc_shard_unpack(self.a.data_ptr(), self.b.data_ptr(), self.n_local, g, lp, self.ctx.stream))
        self.a, self.b = self.b, self.a

    def expect_z_partial(self):
        capi, lib, torch = self.capi, self.capi.load(), self.torch
        out = self.ctx.empty(self.n_local, torch.float64)
        capi.check(lib.dtc_expect_z(self.a.data_ptr(), self.n_local, 1, None, out.data_ptr(), self.ctx.stream))
        nrm = self.ctx.empty(1, torch.float64)
        capi.check(lib.dtc_probs(self.a.data_ptr(), self.n_local, 1, 0, None, None, nrm.data_ptr(), self.ctx.stream))
        return out.cpu().numpy(), float(nrm.cpu().numpy()[0])
