"""Statevector sharded on its top log2(P) qubits across P ranks (BASELINE config C5).

The reference never simulates more than 21 qubits (its snake layout has 21 entries, fast.py:177); this
module runs the same kicked-Ising circuits (dtc_qasm.py:70-91 shape, or any circuit of the supported gate
set) at n = 34-35, where one state is 256-512 GiB.

Layout: global basis index = (rank << n_local) | local index.  Per layer R_j D_j:
  * diagonal terms need no communication -- the rank bits enter the phases (`rank_bits` of the C ABI);
  * rotations act on local qubits only, so the g qubits that are currently global are exchanged with the top g local
    qubits once per layer, and the new qubit->bit permutation is kept (no swap back).
Schedule of one period (steady state; X = the qubits that were global, now at the top local bits):
    T   one sweep of the whole shard on the top tile group:  R_j|X -> D_j -> R_{j+1}|top group        (look-ahead)
    S   slice by slice (slice d = the amplitudes whose top g local bits spell d = what rank d receives):
        the sweeps of R_{j+1} on the other local qubits, then slice d is pushed to rank d over NVLink on a side
        stream while the next slice is swept -- the exchange of layer j+1 hides behind these sweeps.
Noise: the trajectory's Paulis are sampled on the host with the same Philox contract as the device
(`philox_uniform`), and the Pauli frame is tracked on the host while the circuit is cut into segments, so each
segment is an ideal program with sign-resolved angles |theta'| <= pi/2.

The local work goes through an engine object (`CudaShardEngine` binds libdtcsim through capi;
tests/test_sharded_cpu.py plugs a numpy engine to check this host logic with gloo on CPU).
"""
import math

import numpy as np

from . import plan

PI = math.pi
TOP_GROUP = 5      # qubits of a high-stride tile group (mode C of k_tile_stream: 7 passive low bits + 5 qubits)
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
        sliced = hasattr(self.engine, "exchange_sliced") and g > 0
        top = max(g, min(TOP_GROUP, nl - g))       # bits [nl - top, nl): the tile group the exchanged qubits land in
        while j < M:
            rot, d1, d2 = layers[j]
            todo = {q: th for q, th in rot.items() if q not in applied}
            glob = [q for q in todo if phys[q] >= nl]
            if glob:
                loc = {q: th for q, th in todo.items() if phys[q] < nl}
                lq = list(range(nl - g, nl))       # the top g local bits go out
                if sliced:
                    # rotations still due on the outgoing bits (first layer of a run): one sweep of the whole shard on
                    # the top group; everything below then rides with the exchange, slice by slice
                    if any(phys[q] >= nl - g for q in loc):
                        grp = {q: th for q, th in loc.items() if phys[q] >= nl - top}
                        prog = _SegmentProgram(n, nl)
                        prog.add_layer_rot(1, grp, phys)
                        self.engine.run_segment(prog, first)
                        first = False
                        self.stats["segments"] += 1
                        applied |= set(grp)
                        loc = {q: th for q, th in loc.items() if q not in grp}
                    prog = None
                    if loc:
                        prog = _SegmentProgram(n, nl - g)
                        prog.add_layer_rot(1, loc, phys)
                        self.stats["segments"] += 1
                    self.engine.exchange_sliced(prog, lq, first)
                    first = False
                    applied |= set(loc)
                else:
                    if loc:
                        prog = _SegmentProgram(n, nl)
                        prog.add_layer_rot(1, loc, phys)
                        self.engine.run_segment(prog, first)
                        first = False
                        self.stats["segments"] += 1
                        applied |= set(loc)
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
                if sliced and any(phys[q] >= nl for q in layers[j + 1][0]):
                    # an exchange follows: look ahead on the top group only (this segment stays ONE sweep);
                    # the other local rotations of layer j+1 are swept slice by slice under the exchange
                    nxt = {q: th for q, th in nxt.items() if phys[q] >= nl - top}
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


class ThreadFabric:
    """P emulated ranks inside ONE process on ONE GPU (one thread per rank): lets the `-m gpu` tests drive the sharded
    path -- rank bits, slice programs, the exchange permutation -- through the C ABI without a multi-GPU box.  Every
    rank registers its two buffers; an exchange is device copies between them, fenced by thread barriers."""

    def __init__(self, world):
        import threading
        self.world = world
        self.bar = threading.Barrier(world, timeout=300)
        self.bufs = [None] * world          # per rank: {"a": tensor, "b": tensor}

    def register(self, rank, a, b):
        self.bufs[rank] = {"a": a, "b": b}
        self.bar.wait()

    def all_reduce(self, rank, arr):
        if not hasattr(self, "_red"):
            self._red = [None] * self.world
        self._red[rank] = np.asarray(arr, dtype=np.float64).copy()
        self.bar.wait()
        tot = sum(self._red)
        self.bar.wait()
        return tot


class CudaShardEngine:
    """Local shard on one GPU: libdtcsim for the fused sweeps; the exchange is
       * `symm`   (default for world > 1): both state buffers live in symmetric memory (torch.distributed._symmetric_memory:
                  every rank maps every peer's buffers), a finished slice is pushed into the peer's receive buffer by the
                  copy engines over NVLink on a side stream while the SMs sweep the next slice; one device-side barrier
                  per exchange;
       * `nccl`   fallback: pairwise ncclSend/ncclRecv per slice on the side stream (k_tile_stream then leaves a few SMs
                  to the NCCL kernels), or a single all_to_all_single when `overlap` is off;
       * `thread` ranks emulated by threads on one GPU (ThreadFabric; tests)."""

    NCCL_SMS = 20            # SMs left to the NCCL send/recv kernels while sweeps and exchange overlap (nccl mode)
    REMOTE_CTAS = int(__import__("os").environ.get("DTCSIM_REMOTE_CTAS", "64"))   # persistent CTAs of a sweep that stores to a peer
    CE_QUARTERS = int(__import__("os").environ.get("DTCSIM_CE_QUARTERS", "0"))    # quarters of a slice pushed by the copy engine instead
    # (L = 34 on 8 B200, two stores per SM in flight: 24 / 32 / 48 / 56 / 64 CTAs -> 10.6 / 12.6 / 14.8 / 14.4 / 15.1 periods/s;
    #  copy-engine pushes instead: 13.7; profiles/sharded_L34_P8_r2_variants.jsonl)

    def __init__(self, n, n_local, rank, world, device_index, group=None, transport=None, overlap=True, fabric=None):
        import torch
        from . import backend, capi
        self.torch, self.capi = torch, capi
        self.n, self.n_local, self.rank, self.world, self.group = n, n_local, rank, world, group
        self.ctx = backend.DeviceContext(device_index)
        self.overlap = bool(overlap)
        self.fabric = fabric
        if transport is None:
            transport = "thread" if fabric is not None else ("symm" if world > 1 else "none")
        self.transport = transport
        self.hdl = {}
        numel = 1 << n_local
        c128 = torch.complex128
        if transport == "symm":
            try:
                import torch.distributed as dist
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                bufs = []
                for _ in range(2):
                    t = symm_mem.empty(2 * numel, dtype=torch.float64, device=self.ctx.device)
                    h = symm_mem.rendezvous(t, grp)
                    bufs.append((t, h))
                self.a, self.b = bufs[0][0].view(c128), bufs[1][0].view(c128)
                self.hdl = {self.a.data_ptr(): bufs[0][1], self.b.data_ptr(): bufs[1][1]}
            except Exception as exc:                      # no peer mapping on this box: NCCL send/recv instead
                self.transport = transport = "nccl"
                self.symm_error = repr(exc)
        if transport != "symm":
            self.a = self.ctx.empty(numel, c128)
            self.b = self.ctx.empty(numel, c128)
        if transport == "thread":
            fabric.register(rank, self.a, self.b)
        self.comm = torch.cuda.Stream(device=self.ctx.index) if world > 1 else None
        self.ce = torch.cuda.Stream(device=self.ctx.index) if world > 1 else None
        # symm / thread transports: the LAST sweep of a slice program stores its tiles straight into the receiver's buffer
        # (TMA stores to peer memory over NVLink) -- sweep and exchange are one kernel; it runs on the side stream with
        # REMOTE_CTAS persistent CTAs while the other SMs already sweep the next slice
        self.fuse_store = __import__("os").environ.get("DTCSIM_FUSE_STORE", "1") == "1"
        self.fused_stores = 0
        self.n_sms = torch.cuda.get_device_properties(self.ctx.index).multi_processor_count
        self._handles = {}
        self.passes = 0
        self.passes_weighted = 0.0  # state sweeps in units of the whole shard (a slice sweep counts 1 / 2^g)
        self.fast_exchanges = 0
        self.sliced_exchanges = 0
        self.timing = None          # set to {} to collect wall-clock seconds per component (synchronises after each)

    # ---- programs: one handle (+ workspace) per distinct segment, kept for the life of the engine
    def _handle(self, prog, n_local, n_traj=1):
        ev = prog.arrays()
        key = (n_local, n_traj, prog.n_layers) + tuple(ev[k].tobytes() for k in ("type", "layer", "q0", "q1", "slot", "val"))
        hit = self._handles.get(key)
        if hit is None:
            capi = self.capi
            h = capi.ProgramHandle(prog, self.ctx.index, capi.ENGINE_AUTO, n_local)
            wsb = h.workspace_bytes(n_traj)
            hit = (h, self.ctx.empty(wsb, self.torch.uint8), wsb)
            # segment programs are ideal (signs resolved on the host): their sign masks are all zero, written once
            h.prepare(n_traj, 0, 0, hit[1].data_ptr(), wsb, self.ctx.stream)
            if len(self._handles) >= 256:
                old = self._handles.pop(next(iter(self._handles)))
                old[0].close()                            # tables are freed in stream order after their last use
            self._handles[key] = hit
        return hit

    def close(self):
        for h, _ws, _b in self._handles.values():
            h.close()
        self._handles = {}

    def _timed(self, key, t0):
        if self.timing is not None:
            import time
            self.torch.cuda.synchronize(self.ctx.index)
            self.timing[key] = self.timing.get(key, 0.0) + time.perf_counter() - t0

    def _run(self, prog, ptr, n_local, init, rank_bits):
        h, ws, wsb = self._handle(prog, n_local)
        h.run(ptr, 1, 0, 0, ws.data_ptr(), wsb, self.ctx.stream, init_index=init, rank_bits=rank_bits)
        self.passes_weighted += h.num_passes * (1 << n_local) / float(1 << self.n_local)
        return h.num_passes

    def run_segment(self, prog, first):
        import time
        capi = self.capi
        t0 = time.perf_counter()
        init = capi.INIT_KEEP if not first else (0 if self.rank == 0 else capi.INIT_ZERO)
        self.passes += self._run(prog, self.a.data_ptr(), self.n_local, init, self.rank)
        self._timed("segment_sweeps", t0)

    # ---- exchange
    def exchange(self, lq):
        """Swap the g global qubits with the local bits lq (whole shard at once, no overlap)."""
        import time
        capi, lib = self.capi, self.capi.load()
        g = len(lq)
        t0 = time.perf_counter()
        if list(lq) == list(range(self.n_local - g, self.n_local)):
            # the outgoing qubits are the top local bits: chunk d of the state IS what rank d receives and the incoming
            # chunks land where they belong, so pack and unpack are identities -- no extra sweeps
            self.exchange_sliced(None, lq, False)
            self._timed("all_to_all", t0)
            return
        import torch.distributed as dist
        f64 = self.torch.float64
        _, lp = capi.i32(lq)
        capi.check(lib.dtc_shard_pack(self.a.data_ptr(), self.b.data_ptr(), self.n_local, g, lp, self.ctx.stream))
        if self.transport == "thread":
            raise ValueError("ThreadFabric exchanges the top local bits only")
        tmp = self.ctx.empty(1 << self.n_local, self.torch.complex128)
        dist.all_to_all_single(tmp.view(f64), self.b.view(f64), group=self.group)
        capi.check(lib.dtc_shard_unpack(tmp.data_ptr(), self.b.data_ptr(), self.n_local, g, lp, self.ctx.stream))
        self._swap()
        self._timed("all_to_all", t0)

    def _swap(self):
        self.a, self.b = self.b, self.a

    def exchange_sliced(self, prog, lq, first):
        """Exchange of the top g local bits with the g global bits, fused with `prog` (rotations on the lower local
        qubits, or None): slice d -- the part of the shard rank d will own -- is swept and then sent while the next
        slice is swept.  Remote slices go first (staggered so that every rank sends to and receives from one peer
        at a time), the slice that stays on this rank last."""
        torch, capi = self.torch, self.capi
        g, P, r = len(lq), self.world, self.rank
        assert list(lq) == list(range(self.n_local - g, self.n_local)) and (1 << g) == P
        nls = self.n_local - g
        S = 1 << nls
        cur = torch.cuda.current_stream(self.ctx.index)
        a, b = self.a, self.b
        self.sliced_exchanges += 1
        self.fast_exchanges += 1

        def sweep(d):
            if prog is None:
                if first:
                    a[d * S:(d + 1) * S].zero_()
                    if r == 0 and d == 0:
                        a[0:1].fill_(1.0)
                return
            init = capi.INIT_KEEP if not first else (0 if (r == 0 and d == 0) else capi.INIT_ZERO)
            self._run(prog, a.data_ptr() + 16 * d * S, nls, init, (r << g) | d)

        fuse = False
        if prog is not None:
            h, ws, wsb = self._handle(prog, nls)
            np_ = h.num_passes
            self.passes += np_
            fuse = self.fuse_store and not first and self.transport in ("symm", "thread") and self.overlap and nls >= 12
            if fuse:
                try:
                    fuse = h.pass_info(np_ - 1) == (True, True)      # streaming pass with contiguous tiles
                except ValueError:
                    fuse = False

        def sweep_and_send(d, dst_ptr, stream_last, ctas_main, ctas_last, dst_tensor=None):
            """Slice d: all sweeps but the last in place, the last one storing its tiles at dst_ptr (a peer's receive slot);
            with CE_QUARTERS = k the last k quarters of the slice are swept in place and copied instead (dst_tensor)."""
            src = a.data_ptr() + 16 * d * S
            rb = (r << g) | d
            if np_ > 1:
                h.run_passes(src, 0, np_ - 1, 1, ws.data_ptr(), wsb, self.ctx.stream, n_ctas=ctas_main, rank_bits=rb)
            kq = self.CE_QUARTERS if (dst_tensor is not None and nls - 2 >= 12 and max(prog.ev["q0"]) < nls - 2) else 0
            if kq:
                hq, wsq, wsqb = self._handle(prog, nls - 2, n_traj=4)
                Q = S >> 2
                hq.run_passes(src, np_ - 1, np_, 4 - kq, wsq.data_ptr(), wsqb, stream_last, store_last=dst_ptr, n_ctas=ctas_last,
                              rank_bits=rb << 2)
                lo = d * S + (4 - kq) * Q
                hq.run_passes(a.data_ptr() + 16 * lo, np_ - 1, np_, kq, wsq.data_ptr(), wsqb, self.ctx.stream, n_ctas=ctas_main,
                              rank_bits=rb << 2)
                dst_tensor[r * S + (4 - kq) * Q:(r + 1) * S].copy_(a[lo:lo + kq * Q])
            else:
                h.run_passes(src, np_ - 1, np_, 1, ws.data_ptr(), wsb, stream_last, store_last=dst_ptr, n_ctas=ctas_last,
                             rank_bits=rb)
            self.passes_weighted += np_ * (1 << nls) / float(1 << self.n_local)
            self.fused_stores += 1

        if self.transport == "thread":
            # emulated ranks on one GPU: the "peer" buffers are the other threads' tensors
            torch.cuda.synchronize(self.ctx.index)
            self.fabric.bar.wait()                     # every rank's receive buffer is free
            if fuse:
                for d in range(P):
                    sweep_and_send(d, self.fabric.bufs[d]["b"].data_ptr() + 16 * r * S, self.ctx.stream, 0, 0,
                                   dst_tensor=self.fabric.bufs[d]["b"])
                torch.cuda.synchronize(self.ctx.index)
                self.fabric.bar.wait()
            else:
                for d in range(P):
                    sweep(d)
                torch.cuda.synchronize(self.ctx.index)
                self.fabric.bar.wait()
                for src in range(P):
                    b[src * S:(src + 1) * S].copy_(self.fabric.bufs[src]["a"][r * S:(r + 1) * S])
                torch.cuda.synchronize(self.ctx.index)
                self.fabric.bar.wait()
            self._swap()
            self.fabric.bufs[r] = {"a": self.a, "b": self.b}
            self.fabric.bar.wait()
            return

        if not self.overlap or self.transport == "none":
            for d in range(P):
                sweep(d)
            if P > 1:
                import torch.distributed as dist
                f64 = torch.float64
                dist.all_to_all_single(b.view(f64), a.view(f64), group=self.group)
                self._swap()
            return

        order = [(r + s) % P for s in range(1, P)] + [r]
        comm = self.comm
        if self.transport == "symm" and fuse:
            # fused sweep + exchange: the last sweep of slice d writes rank d's receive slot directly (peer memory, NVLink);
            # it runs on the side stream on REMOTE_CTAS SMs while the remaining SMs sweep the next slice
            import ctypes
            hb = self.hdl[b.data_ptr()]
            peers = hb.buffer_ptrs
            comm_ptr = ctypes.c_void_p(comm.cuda_stream)
            remote = self.REMOTE_CTAS
            if P == 2 and "DTCSIM_REMOTE_CTAS" not in __import__("os").environ:
                remote = 32                  # one remote slice only: measured 61 ms per period at L = 32 (CE pushes: 73 ms)
            n_last = max(1, min(remote, self.n_sms - 1))
            n_main = self.n_sms - n_last
            # Hybrid (CE_QUARTERS = k > 0): the peer-storing sweep gets ~12 GB/s per SM out of the link, a copy engine 750 GB/s
            # for no SM at all but one more read of the data from HBM.  The slice programs rotate only bits below
            # n_local - g - 2, so the LAST sweep of a slice can run quarter by quarter: 4 - k quarters store into the peer
            # (side stream, REMOTE_CTAS SMs), k quarters are swept in place by the main SMs and pushed by the copy engine.
            kq = self.CE_QUARTERS if (nls - 2 >= 12 and max(prog.ev["q0"]) < nls - 2) else 0
            if kq:
                hq, wsq, wsqb = self._handle(prog, nls - 2, n_traj=4)
                Q = S >> 2
                ce = self.ce
            for d in order:
                src = a.data_ptr() + 16 * d * S
                rb = (r << g) | d
                if np_ > 1:
                    h.run_passes(src, 0, np_ - 1, 1, ws.data_ptr(), wsb, self.ctx.stream, n_ctas=n_main, rank_bits=rb)
                if d != r:
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    comm.wait_event(ev)
                if d != r and kq:
                    hq.run_passes(src, np_ - 1, np_, 4 - kq, wsq.data_ptr(), wsqb, comm_ptr,
                                  store_last=int(peers[d]) + 16 * r * S, n_ctas=n_last, rank_bits=rb << 2)
                    lo = d * S + (4 - kq) * Q                      # the last kq quarters: in place, then the copy engine
                    hq.run_passes(a.data_ptr() + 16 * lo, np_ - 1, np_, kq, wsq.data_ptr(), wsqb, self.ctx.stream, n_ctas=n_main,
                                  rank_bits=rb << 2)
                    ev2 = torch.cuda.Event()
                    ev2.record(cur)
                    ce.wait_event(ev2)
                    with torch.cuda.stream(ce):
                        peer_b = hb.get_buffer(d, (2 * kq * Q,), torch.float64, 2 * (r * S + (4 - kq) * Q))
                        peer_b.copy_(a[lo:lo + kq * Q].view(torch.float64), non_blocking=True)
                elif d != r:
                    h.run_passes(src, np_ - 1, np_, 1, ws.data_ptr(), wsb, comm_ptr, store_last=int(peers[d]) + 16 * r * S,
                                 n_ctas=n_last, rank_bits=rb)
                else:                            # the slice that stays: stored into this rank's own receive buffer, main stream
                    h.run_passes(src, np_ - 1, np_, 1, ws.data_ptr(), wsb, self.ctx.stream, store_last=b.data_ptr() + 16 * r * S,
                                 n_ctas=n_main, rank_bits=rb)
                self.passes_weighted += np_ * (1 << nls) / float(1 << self.n_local)
                self.fused_stores += 1
            comm.wait_stream(cur)
            if kq:
                comm.wait_stream(ce)
            with torch.cuda.stream(comm):
                hb.barrier(channel=0)            # every rank's stores into every receive buffer have completed
            cur.wait_stream(comm)
            self._swap()
            return

        if self.transport == "symm":
            hb = self.hdl[b.data_ptr()]
            for d in order:
                sweep(d)
                if d == r:
                    b[r * S:(r + 1) * S].copy_(a[r * S:(r + 1) * S], non_blocking=True)
                    continue
                ev = torch.cuda.Event()
                ev.record(cur)
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    peer_b = hb.get_buffer(d, (2 * S,), torch.float64, 2 * r * S)     # slot r of rank d's receive buffer
                    peer_b.copy_(a[d * S:(d + 1) * S].view(torch.float64), non_blocking=True)
            comm.wait_stream(cur)                # ... and this rank has finished reading its send buffer
            with torch.cuda.stream(comm):
                hb.barrier(channel=0)            # every push into every receive buffer has completed
            cur.wait_stream(comm)
            self._swap()
            return

        # nccl: pairwise send / recv per step on the side stream; the sweeps leave NCCL_SMS SMs to those kernels
        import torch.distributed as dist
        f64 = torch.float64
        capi.set_stream_ctas(max(1, torch.cuda.get_device_properties(self.ctx.index).multi_processor_count - self.NCCL_SMS))
        try:
            works = []
            for s_, d in enumerate(order):
                sweep(d)
                if d == r:
                    b[r * S:(r + 1) * S].copy_(a[r * S:(r + 1) * S], non_blocking=True)
                    continue
                src = (r - (s_ + 1)) % P
                ev = torch.cuda.Event()
                ev.record(cur)
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    ops = [dist.P2POp(dist.isend, a[d * S:(d + 1) * S].view(f64), d, group=self.group),
                           dist.P2POp(dist.irecv, b[src * S:(src + 1) * S].view(f64), src, group=self.group)]
                    works += dist.batch_isend_irecv(ops)
            with torch.cuda.stream(comm):
                for w in works:
                    w.wait()
            cur.wait_stream(comm)
        finally:
            capi.set_stream_ctas(0)
        self._swap()

    def expect_z_partial(self):
        capi, lib, torch = self.capi, self.capi.load(), self.torch
        out = self.ctx.empty(self.n_local, torch.float64)
        capi.check(lib.dtc_expect_z(self.a.data_ptr(), self.n_local, 1, None, out.data_ptr(), self.ctx.stream))
        nrm = self.ctx.empty(1, torch.float64)
        capi.check(lib.dtc_probs(self.a.data_ptr(), self.n_local, 1, 0, None, None, nrm.data_ptr(), self.ctx.stream))
        return out.cpu().numpy(), float(nrm.cpu().numpy()[0])
