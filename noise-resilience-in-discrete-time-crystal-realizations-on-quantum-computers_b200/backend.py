"""AerSimulator-compatible backend: the Python host of the drop-in boundary.

Mirrors exactly the surface the reference scripts touch:

    backend = AerSimulator(noise_model=noise_model, device="GPU", cuStateVec_enable=True)   # fast.py:156
    passmanager = generate_preset_pass_manager(optimization_level=0, backend=backend, ...)  # fast.py:181
    backend_name = getattr(backend, 'name', ...)                                            # fast.py:191
    result = backend.run(circ_tnoise, shots=1024).result()                                  # fast.py:211
    counts = result.get_counts(circ_tnoise)                                                 # fast.py:212

``run`` compiles each circuit with plan.compile_circuit and executes it on the CUDA library through
capi (ctypes).  torch only provides device buffers and the current stream.  Method selection follows
Aer's ``automatic`` rule: exact density matrix when noise is present and shots > 2^n,
otherwise one Pauli trajectory per shot (noisy) or a single statevector sampled `shots` times (ideal).
Unsupported input raises ValueError; a missing CUDA library or GPU raises RuntimeError.
"""
import ctypes
import os
import time

import numpy as np

from . import capi
from .ir import as_circuit
from .noise import as_noise_model
from .plan import compile_circuit
from . import philox_np

MAX_DM_QUBITS = 13
PIPELINE_DEPTH = 8        # circuits of a run(list) the host may enqueue ahead of the one whose results it fetches
MAX_PROB_QUBITS = 12


def compute_z_expectation(counts, num_qubits):
    """Same reducer as the reference (fast.py:92-109), provided for convenience."""
    total = sum(counts.values())
    out = []
    for k in range(num_qubits):
        p0 = sum(c for b, c in counts.items() if b[::-1][k] == "0")
        out.append((p0 - (total - p0)) / total)
    return out


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("dtcsim needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    return torch


class DeviceContext:
    """Device buffers + stream for one GPU (torch is plumbing only)."""

    def __init__(self, index=None):
        torch = _torch()
        self.torch = torch
        self.index = torch.cuda.current_device() if index is None else int(index)
        self.device = torch.device("cuda", self.index)
        capi.load()

    @property
    def stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.index).cuda_stream)

    def empty(self, n, dtype):
        return self.torch.empty(int(n), dtype=dtype, device=self.device)

    def free_bytes(self):
        free, _ = self.torch.cuda.mem_get_info(self.index)
        return int(free)

    def sync(self):
        self.torch.cuda.synchronize(self.index)


class TrajectoryBatch:
    """Result of evolving a batch of trajectories: psi' on the device plus their Pauli frames."""

    def __init__(self, ctx, handle, state, ws, n_traj, traj_offset):
        self.ctx, self.handle, self.state, self.ws = ctx, handle, state, ws
        self.n_traj, self.traj_offset = n_traj, traj_offset
        self.n = handle.n_local
        self.fx, self.fz, self.ph = handle.frames(ws.data_ptr(), n_traj)
        self.fused_rdm = None           # set by evolve(fused_rdm=True): [n_traj, 2, 2] view into the workspace

    def probs(self, qubits, apply_frame=True):
        """[n_traj, 2^k] outcome probabilities (torch float64 on device)."""
        torch = self.ctx.torch
        k = len(qubits)
        out = self.ctx.empty(self.n_traj << k, torch.float64)
        _, qp = capi.i32(qubits)
        capi.check(capi.load().dtc_probs(self.state.data_ptr(), self.n, self.n_traj, k, qp,
                                         self.fx if apply_frame else None, out.data_ptr(), self.ctx.stream))
        return out.view(self.n_traj, 1 << k)

    def rdm(self, bits):
        """[n_traj, 2^k, 2^k] reduced density matrices of psi' on `bits` (k <= 2), torch complex128."""
        torch = self.ctx.torch
        k = len(bits)
        out = self.ctx.empty(self.n_traj << (2 * k), torch.complex128)
        _, qp = capi.i32(list(bits) if k else [0])
        capi.check(capi.load().dtc_rdm(self.state.data_ptr(), self.n, self.n_traj, k, qp, out.data_ptr(),
                                       self.ctx.stream))
        return out.view(self.n_traj, 1 << k, 1 << k)

    def readout_rdm(self):
        """Reduced density matrix the read-out of a factorised circuit starts from: the fused one of the last pass
        if evolve() ran with fused_rdm=True, else a dtc_rdm reduction of the state."""
        if self.fused_rdm is not None:
            return self.fused_rdm
        return self.rdm(self.handle.prog.small["reg_bits"])

    def masks_host(self, first_layer=0):
        """Sign masks [n_layers - first_layer, 4, n_traj] (uint64) written by the device frame walk."""
        torch = self.ctx.torch
        nl, n = self.handle.prog.n_layers, self.n_traj
        m = self.ws[: nl * 4 * n * 8].view(torch.int64).view(nl, 4, n)[first_layer:]
        return m.cpu().numpy().view(np.uint64)

    def outcome_probs(self, rdm=None):
        """Per-trajectory probabilities of the measured qubits [n_traj, 2^m] (frame applied), on the device.
        Uses the read-out factorisation when the program was compiled with it (rdm: result of an earlier
        self.rdm(reg_bits) call, so the state buffer may already have been reused)."""
        prog = self.handle.prog
        torch = self.ctx.torch
        if prog.small is None:
            return self.probs([q for q, _ in prog.measures])
        if rdm is None:
            rdm = self.readout_rdm()
        out = self.ctx.empty(self.n_traj << len(prog.measures), torch.float64)
        self.handle.readout(rdm.data_ptr(), self.ws.data_ptr(), self.n_traj, out.data_ptr(), self.ctx.stream)
        return out.view(self.n_traj, 1 << len(prog.measures))

    def expect_z(self, apply_frame=True):
        torch = self.ctx.torch
        out = self.ctx.empty(self.n_traj * self.n, torch.float64)
        capi.check(capi.load().dtc_expect_z(self.state.data_ptr(), self.n, self.n_traj,
                                            self.fx if apply_frame else None, out.data_ptr(), self.ctx.stream))
        return out.view(self.n_traj, self.n)

    def sample_states(self, seed, n_samples=1):
        """Basis-state samples from |psi|^2 (frame applied): [n_traj] (n_samples == 1) or [n_traj, n_samples] int64."""
        torch = self.ctx.torch
        nchunk = 1 << max(self.n - 12, 0)
        scratch = self.ctx.empty(self.n_traj * nchunk, torch.float64)
        out = self.ctx.empty(self.n_traj * n_samples, torch.int64)
        capi.check(capi.load().dtc_sample_states_multi(self.state.data_ptr(), self.n, self.n_traj, int(n_samples),
                                                       int(seed) & 0xFFFFFFFFFFFFFFFF, self.traj_offset, self.fx,
                                                       scratch.data_ptr(), out.data_ptr(), self.ctx.stream))
        return out if n_samples == 1 else out.view(self.n_traj, n_samples)

    def materialize(self):
        """Apply the frames in place: the buffer then holds the true statevectors."""
        capi.check(capi.load().dtc_materialize(self.state.data_ptr(), self.n, self.n_traj, self.fx, self.fz,
                                               self.ph, None, self.ctx.stream))
        return self.state.view(self.n_traj, 1 << self.n)

    def frames_host(self):
        torch = self.ctx.torch
        self.ctx.sync()
        n = self.n_traj
        nl4 = self.handle.prog.n_layers * 4
        w = self.ws
        u64 = w[: (nl4 + 2) * n * 8].view(torch.int64)
        fx = u64[nl4 * n:(nl4 + 1) * n].cpu().numpy().view(np.uint64)
        fz = u64[(nl4 + 1) * n:(nl4 + 2) * n].cpu().numpy().view(np.uint64)
        ph = w[(nl4 + 2) * n * 8:(nl4 + 2) * n * 8 + 4 * n].view(torch.int32).cpu().numpy()
        return fx, fz, ph


def evolve(ctx, prog, n_traj, traj_offset=0, seed=0, engine=capi.ENGINE_AUTO, handle=None, state=None,
           init_index=0, fused_rdm=False):
    """Run the compiled program for a batch of trajectories; returns a TrajectoryBatch.

    fused_rdm=True (read-out only: the caller will not look at the state): where the library can, the last pass
    reduces the read-out qubit's density matrix instead of storing the state (batch.fused_rdm)."""
    torch = ctx.torch
    own = handle is None
    if own:
        handle = capi.ProgramHandle(prog, ctx.index, engine)
    n = handle.n_local
    need = n_traj << n
    wsb = handle.workspace_bytes(n_traj)
    ws = ctx.empty(wsb, torch.uint8)
    fused = handle.set_fused_rdm(bool(fused_rdm) and getattr(prog, "small", None) is not None)
    resident, group, scratch_bytes = handle.resident_info(n_traj) if (fused and capi.RESIDENT) else (False, 0, 0)
    if resident:
        # read-out-only run of a program whose passes all stream: one persistent launch, `group` L2-resident state slots
        need = scratch_bytes // 16
        if state is None or state.numel() < need:
            state = ctx.empty(need, torch.complex128)
        handle.run_resident(state.data_ptr(), scratch_bytes, n_traj, traj_offset, seed, ws.data_ptr(), wsb, ctx.stream,
                            init_index=init_index)
    else:
        if state is None or state.numel() < need:
            state = ctx.empty(need, torch.complex128)
        handle.run(state.data_ptr(), n_traj, traj_offset, seed, ws.data_ptr(), wsb, ctx.stream, init_index=init_index)
    batch = TrajectoryBatch(ctx, handle, state[:need], ws, n_traj, traj_offset)
    batch.resident = resident
    if fused:
        off = handle.fused_rdm_ptr(ws.data_ptr(), n_traj) - ws.data_ptr()
        batch.fused_rdm = ws[off:off + 64 * n_traj].view(torch.complex128).view(n_traj, 2, 2)
    return batch


def sample_rows(ctx, probs, n_samples, seed, traj_offset):
    torch = ctx.torch
    rows, cols = probs.shape
    out = ctx.empty(rows * n_samples, torch.int32)
    capi.check(capi.load().dtc_sample_rows(probs.data_ptr(), rows, cols, n_samples,
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, traj_offset, out.data_ptr(), ctx.stream))
    return out.view(rows, n_samples)


def flatten_dm_segments(segments):
    """prog.dm_segments (R / D / N segments only) -> the flat arrays of dtc_dm_run (seg_type, seg_off, q0, q1, val, probs)."""
    seg_type, seg_off, q0, q1, val, probs = [], [0], [], [], [], []
    for seg in segments:
        if seg[0] == "R":
            seg_type.append(0)
            for q, th in seg[1]:
                q0.append(q); q1.append(-1); val.append(th); probs.append((0.0, 0.0, 0.0))
        elif seg[0] == "D":
            seg_type.append(1)
            for q, a in seg[1].items():
                q0.append(q); q1.append(-1); val.append(a); probs.append((0.0, 0.0, 0.0))
            for (i, j), b in seg[2].items():
                q0.append(i); q1.append(j); val.append(b); probs.append((0.0, 0.0, 0.0))
        elif seg[0] == "N":
            seg_type.append(2)
            for q, pr in seg[1]:
                q0.append(q); q1.append(-1); val.append(0.0); probs.append(tuple(pr))
        else:
            raise ValueError(f"segment type {seg[0]!r} does not go through dtc_dm_run")
        seg_off.append(len(q0))
    return (np.asarray(seg_type, dtype=np.int32), np.asarray(seg_off, dtype=np.int32), np.asarray(q0, dtype=np.int32),
            np.asarray(q1, dtype=np.int32), np.asarray(val, dtype=np.float64),
            np.asarray(probs, dtype=np.float64).reshape(-1, 3))


def split_dm_segments(segments):
    """Runs of R / D / N segments (one dtc_dm_run call each) separated by the "K" segments of non-Pauli channels
    (one dtc_dm_superop call per channel): [("run", [segments]) | ("K", [(qubit, superoperator), ...]), ...]."""
    out, cur = [], []
    for seg in segments:
        if seg[0] == "K":
            if cur:
                out.append(("run", cur))
                cur = []
            out.append(("K", list(seg[1])))
        else:
            cur.append(seg)
    if cur:
        out.append(("run", cur))
    return out


def run_density_matrix(ctx, prog, stats=None):
    """Exact noisy evolution of rho (Aer method density_matrix); returns the rho tensor (2^n x 2^n, [col,row]).
    One C-ABI call for the whole program (dtc_dm_run) -- or one per stretch between non-Pauli channels, each of which is one
    sweep of dtc_dm_superop; stats["sweeps"] = passes over rho made."""
    torch = ctx.torch
    lib = capi.load()
    n = prog.n
    if n > MAX_DM_QUBITS:
        raise ValueError(f"density-matrix method supports at most {MAX_DM_QUBITS} qubits (got {n})")
    rho = ctx.empty(1 << (2 * n), torch.complex128)
    s = ctx.stream
    capi.check(lib.dtc_dm_init(rho.data_ptr(), n, 0, s))
    total = 0
    for kind, payload in split_dm_segments(prog.dm_segments):
        if kind == "K":
            for q, S in payload:
                Sf = np.ascontiguousarray(np.asarray(S, dtype=np.complex128).reshape(16)).view(np.float64)
                capi.check(lib.dtc_dm_superop(rho.data_ptr(), n, int(q), Sf.ctypes.data_as(capi.c_f64p), s))
                total += 1
            continue
        st, so, q0, q1, val, pr = flatten_dm_segments(payload)
        sweeps = ctypes.c_int(0)
        capi.check(lib.dtc_dm_run(rho.data_ptr(), n, len(st), st.ctypes.data_as(capi.c_i32p), so.ctypes.data_as(capi.c_i32p),
                                  q0.ctypes.data_as(capi.c_i32p), q1.ctypes.data_as(capi.c_i32p),
                                  val.ctypes.data_as(capi.c_f64p), pr.ctypes.data_as(capi.c_f64p), ctypes.byref(sweeps), s))
        total += sweeps.value
    if stats is not None:
        stats["sweeps"] = total
    return rho


class ExperimentResult:
    def __init__(self, name, counts, data, shots, seed):
        self.name = name
        self.counts = counts
        self.data = data
        self.shots = shots
        self.seed_simulator = seed
        self.success = True


class Result:
    """Subset of qiskit.result.Result the reference consumes (fast.py:211-212)."""

    def __init__(self, experiments, circuits, single, backend_name, time_taken):
        self.results = experiments
        self._circuits = circuits
        self._single = single
        self.backend_name = backend_name
        self.time_taken = time_taken
        self.success = True

    def _index(self, key):
        if key is None:
            return None
        if isinstance(key, int):
            return key
        for i, c in enumerate(self._circuits):
            if c is key:
                return i
        name = key if isinstance(key, str) else getattr(key, "name", None)
        for i, e in enumerate(self.results):
            if e.name == name:
                return i
        raise KeyError(f"circuit {key!r} not found in this result")

    def get_counts(self, experiment=None):
        i = self._index(experiment)
        if i is None:
            if self._single:
                return dict(self.results[0].counts)
            return [dict(e.counts) for e in self.results]
        return dict(self.results[i].counts)

    def data(self, experiment=None):
        i = self._index(experiment)
        return self.results[0 if i is None else i].data

    def expectation_z(self, experiment=None):
        """Exact <Z> of each classical bit from the simulated probabilities (no shot noise)."""
        return self.data(experiment)["expval_z"]


class Job:
    def __init__(self, result):
        self._result = result

    def result(self, timeout=None):
        return self._result

    def status(self):
        return "DONE"

    def job_id(self):
        return f"dtcsim-{id(self):x}"


class DTCSimulator:
    """Drop-in for ``qiskit_aer.AerSimulator`` on the reference's hot path."""

    name = "aer_simulator"            # committed gate_counts file names embed this (fast.py:191,196)
    num_qubits = 64                   # physical width offered to the pass manager (>= 31, fast.py:177)
    basis_gates = ["cx", "id", "rz", "sx", "u1", "u2", "u3"]

    def __init__(self, noise_model=None, method="automatic", device="GPU", cuStateVec_enable=False,
                 seed_simulator=None, shots=1024, cuda_device=None, max_memory_bytes=None, engine="auto",
                 optimize=True, **ignored):
        self.noise_model = noise_model
        self.method = method
        self.seed_simulator = seed_simulator
        self.default_shots = shots
        self.cuda_device = cuda_device
        self.max_memory_bytes = max_memory_bytes
        self.optimize = bool(optimize)       # read-out factorisation (plan.compile_circuit(optimize=True))
        self.engine = {"auto": capi.ENGINE_AUTO, "generic": capi.ENGINE_GENERIC, "tile": capi.ENGINE_TILE}[engine]
        self.sim_options = dict(device=device, cuStateVec_enable=cuStateVec_enable, **ignored)   # accepted, unused
        self._ctx = None

    def set_options(self, **kw):
        for k, v in kw.items():
            if k in ("noise_model", "method", "seed_simulator"):
                setattr(self, k, v)
            elif k == "shots":
                self.default_shots = v
            else:
                self.sim_options[k] = v

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = DeviceContext(self.cuda_device)
        return self._ctx

    # ------------------------------------------------------------------------------------------
    def run(self, circuits, shots=None, seed_simulator=None, noise_model=None, method=None, **opts):
        t0 = time.time()
        single = not isinstance(circuits, (list, tuple))
        circ_list = [circuits] if single else list(circuits)
        shots = self.default_shots if shots is None else int(shots)
        if shots < 1:
            raise ValueError("shots must be positive")
        seed = self.seed_simulator if seed_simulator is None else seed_simulator
        if seed is None:
            seed = int.from_bytes(os.urandom(6), "little")
        if isinstance(seed, (list, tuple, np.ndarray)):          # one seed per circuit (sweeps.run_sweep)
            if len(seed) != len(circ_list):
                raise ValueError("seed_simulator list must have one entry per circuit")
            seeds = [int(x) for x in seed]
        else:
            seeds = [int(seed) + i for i in range(len(circ_list))]
        nm = as_noise_model(self.noise_model if noise_model is None else noise_model)
        method = self.method if method is None else method
        # A list of circuits is pipelined: the host compiles and enqueues circuits ahead of the GPU -- up to
        # PIPELINE_DEPTH circuits are in flight before the oldest one's results are fetched (on a side stream), so a slow
        # compile or a host hiccup does not leave the GPU idle.  Counts are identical to one-by-one calls.
        exps = [None] * len(circ_list)
        pending = []
        handles = []
        try:
            for i, c in enumerate(circ_list):
                r = self._run_one(as_circuit(c), shots, seeds[i], nm, method, getattr(c, "name", None), handles)
                if callable(r):
                    pending.append((i, r))
                else:
                    exps[i] = r
                while len(pending) > PIPELINE_DEPTH:
                    j, fetch = pending.pop(0)
                    exps[j] = fetch()
            for j, fetch in pending:
                exps[j] = fetch()
        finally:
            for h in handles:
                h.close()
        return Job(Result(exps, circ_list, single, self.name, time.time() - t0))

    def _choose_method(self, method, n, shots, nm):
        if method in ("automatic", None):
            if nm is not None and shots > (1 << n) and n <= MAX_DM_QUBITS:
                return "density_matrix"
            return "statevector"
        if method not in ("statevector", "density_matrix"):
            raise ValueError(f"unsupported simulation method {method!r}")
        return method

    def _run_one(self, circ, shots, seed, nm, method, name, handles):
        """Returns the ExperimentResult, or (noisy trajectory runs) a callable that fetches it once the enqueued
        GPU work is done; program handles to close after that go to `handles`."""
        t0 = time.time()
        torch = self.ctx.torch
        ctx = self.ctx
        full_nm = nm
        if nm is not None and not nm.has_gate_noise():
            nm = None                      # readout errors only: the evolution is ideal, the recorded bits are not
        channels = nm is not None and nm.has_channel_noise()
        if channels:
            # non-Pauli channels (thermal relaxation, ...) exist only as density-matrix segments: exact rho whatever the
            # shot count (Aer would sample Kraus trajectories for shots <= 2^n; the outcome distribution is the same)
            if method not in ("automatic", None, "density_matrix"):
                raise ValueError("non-Pauli noise channels need method='density_matrix' (or 'automatic')")
            prog0 = self._compiled(circ, nm, want_dm=True)
            if prog0.n > MAX_DM_QUBITS:
                raise ValueError(f"non-Pauli noise channels run on the density-matrix method only: at most {MAX_DM_QUBITS} "
                                 f"active qubits (got {prog0.n})")
            method = "density_matrix"
        else:
            prog0 = self._compiled(circ, nm)
        n = prog0.n
        method = self._choose_method(method, n, shots, nm)
        # classical readout errors (device-calibrated noise, fast.py:77-78): assignment matrix per classical bit
        readout = {}
        if full_nm is not None and full_nm.has_readout_noise():
            for c, q in prog0.measure_orig.items():
                m = full_nm.lookup_readout(q)
                if m is not None:
                    readout[int(c)] = m
        meas = prog0.measures
        if not meas:
            raise ValueError("circuit has no measurements: nothing to count")
        mq = [q for q, _ in meas]
        k = len(mq)
        cbits = np.array([c for _, c in meas], dtype=np.int64)

        def to_clbits(cols):
            cols = np.asarray(cols, dtype=np.int64)
            vals = np.zeros_like(cols)
            for i in range(k):
                vals |= ((cols >> i) & 1) << cbits[i]
            return vals

        data = {"method": method, "n_qubits": n, "register_qubits": prog0.n_main, "active_qubits": prog0.active,
                "seed_simulator": seed}
        if method == "density_matrix":
            # rho is built by its own compile (no read-out factorisation), whose internal bit order differs from
            # prog0's: the measured bits must come from THAT program
            prog = self._compiled(circ, nm, want_dm=True)
            meas = prog.measures
            mq = [q for q, _ in meas]
            cbits = np.array([c for _, c in meas], dtype=np.int64)
            rho = run_density_matrix(ctx, prog)
            probs = ctx.empty(1 << k, torch.float64)
            _, qp = capi.i32(mq)
            capi.check(capi.load().dtc_dm_probs(rho.data_ptr(), n, k, qp, probs.data_ptr(), ctx.stream))
            cols = sample_rows(ctx, probs.view(1, -1), shots, seed, 0).cpu().numpy()[0]
            p_host = probs.cpu().numpy()
            data["probabilities"] = self._clbit_probs(p_host, to_clbits, prog0.n_clbits)
            vals = self._readout_flips(to_clbits(cols), readout, seed, np.arange(shots))
        elif nm is None:
            batch = evolve(ctx, prog0, 1, 0, seed, self.engine)
            data["num_passes"] = batch.handle.num_passes
            if k > MAX_PROB_QUBITS:
                # wide measurement (dtc_qasm.py measure-all at L = 20): shots = basis-state samples of the one state,
                # exact <Z> of every measured qubit from one dtc_expect_z pass (dtc_qasm.py:145)
                idx = batch.sample_states(seed, shots).cpu().numpy().reshape(-1)
                cols = np.zeros(shots, dtype=np.int64)
                for i, q in enumerate(mq):
                    cols |= ((idx >> q) & 1) << i
                ez = batch.expect_z().cpu().numpy()[0]
                data["expval_z_by_clbit"] = {int(c): float(ez[q]) for q, c in meas}
            else:
                probs = batch.outcome_probs()
                cols = sample_rows(ctx, probs, shots, seed, 0).cpu().numpy()[0]
                data["probabilities"] = self._clbit_probs(probs.cpu().numpy()[0], to_clbits, prog0.n_clbits)
            vals = self._readout_flips(to_clbits(cols), readout, seed, np.arange(shots))
        else:
            handle = capi.ProgramHandle(prog0, ctx.index, self.engine)
            handles.append(handle)
            nm_ = prog0.n_main
            per = 16 << nm_
            budget = self.max_memory_bytes or int(0.7 * ctx.free_bytes())
            bt = max(1, min(shots, budget // per))
            state = None
            if k <= MAX_PROB_QUBITS and prog0.small is not None and capi.RESIDENT and handle.set_fused_rdm(True):
                resident, _g, sbytes = handle.resident_info(shots)
                if resident:                              # whole batch in one persistent launch, a few L2-resident state slots
                    bt = shots
                    state = self._state_buffer(sbytes // 16)
            if state is None:
                state = self._state_buffer(bt << nm_)
            queued = []                                   # per batch: (offset, n, device columns / indices, device prob sums)
            ez_sum = None                                 # wide measurements: sum over trajectories of <Z_q> (device)
            for a in range(0, shots, bt):
                nt = min(bt, shots - a)
                batch = evolve(ctx, prog0, nt, a, seed, handle=handle, state=state, fused_rdm=k <= MAX_PROB_QUBITS)
                if k <= MAX_PROB_QUBITS:
                    probs = batch.outcome_probs()
                    queued.append((a, nt, sample_rows(ctx, probs, 1, seed, a), probs.sum(dim=0), batch))
                else:
                    queued.append((a, nt, batch.sample_states(seed), None, batch))
                    ezs = batch.expect_z().sum(dim=0)
                    ez_sum = ezs if ez_sum is None else ez_sum + ezs
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(ctx.index))
            data["num_passes"] = handle.num_passes
            data["trajectories"] = shots

            def finish():
                side = self._side_stream()
                side.wait_event(done)
                vals = np.zeros(shots, dtype=np.int64)
                psum = None
                with torch.cuda.stream(side):
                    for a, nt, cols_dev, ps_dev, _batch in queued:
                        if ps_dev is not None:
                            cols = cols_dev.cpu().numpy()[:, 0]
                            ps = ps_dev.cpu().numpy()
                            psum = ps if psum is None else psum + ps
                        else:
                            idx = cols_dev.cpu().numpy()
                            cols = np.zeros(nt, dtype=np.int64)
                            for i, q in enumerate(mq):
                                cols |= ((idx >> q) & 1) << i
                        vals[a:a + nt] = self._readout_flips(to_clbits(cols), readout, seed, np.arange(a, a + nt))
                if psum is not None:
                    data["probabilities"] = self._clbit_probs(psum / shots, to_clbits, prog0.n_clbits)
                if ez_sum is not None:
                    with torch.cuda.stream(side):
                        ez = ez_sum.cpu().numpy() / shots
                    data["expval_z_by_clbit"] = {int(c): float(ez[q]) for q, c in meas}
                return self._experiment(vals, data, prog0, name or circ.name, shots, seed, t0, readout)

            return finish
        return self._experiment(vals, data, prog0, name or circ.name, shots, seed, t0, readout)

    def _compiled(self, circ, nm, want_dm=False):
        """compile_circuit() with a small cache keyed by the circuit's op list and the noise model: sweeps that run the
        same circuit again (other seeds, other shot counts: shots.py:49) skip the host compile."""
        noise_key = None
        if nm is not None:
            noise_key = (tuple(sorted((k, e.probs) for k, e in nm._all.items())),
                         tuple(sorted((k, e.probs) for k, e in nm._local.items())))
        optimize = self.optimize and not want_dm
        key = (circ.num_qubits, circ.num_clbits, float(circ.global_phase), optimize, want_dm, noise_key,
               tuple(op.astuple() for op in circ.ops))
        cache = self.__dict__.setdefault("_prog_cache", {})
        prog = cache.get(key)
        if prog is None:
            prog = compile_circuit(circ, nm, want_dm=want_dm, optimize=optimize)
            if len(cache) >= 128:
                cache.pop(next(iter(cache)))
            cache[key] = prog
        return prog

    def _state_buffer(self, n_amps):
        """One state buffer per simulator object, grown on demand (stream-ordered reuse across pipelined circuits)."""
        torch = self.ctx.torch
        buf = getattr(self, "_state", None)
        if buf is None or buf.numel() < n_amps:
            self._state = None
            buf = self._state = self.ctx.empty(n_amps, torch.complex128)
        return buf

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = self.ctx.torch.cuda.Stream(device=self.ctx.index)
        return self._side

    @staticmethod
    def _readout_flips(vals, readout, seed, shot_ids):
        """Recorded classical-register values: bit c of shot s flips with P(recorded != true) of its assignment matrix;
        u = philox(seed; index = c, stream 2, trajectory word = shot id) (the test-side restatement draws the same stream)."""
        if not readout:
            return vals
        vals = np.array(vals, dtype=np.int64, copy=True)
        for c, m in readout.items():
            u = philox_np.uniform(seed, c, philox_np.STREAM_READOUT, np.asarray(shot_ids, dtype=np.uint64))
            bit = (vals >> c) & 1
            flip = np.where(bit == 0, u < m[0][1], u < m[1][0])
            vals ^= flip.astype(np.int64) << c
        return vals

    @staticmethod
    def _readout_probs(pr, readout):
        """Exact distribution of the RECORDED register: every classical bit mixed by its assignment matrix."""
        for c, m in readout.items():
            new = {}
            for v, p in pr.items():
                b = (v >> c) & 1
                new[v] = new.get(v, 0.0) + p * m[b][b]
                new[v ^ (1 << c)] = new.get(v ^ (1 << c), 0.0) + p * m[b][1 - b]
            pr = new
        return pr

    @staticmethod
    def _experiment(vals, data, prog0, name, shots, seed, t0, readout=None):
        counts = {}
        uniq, cnt = np.unique(vals, return_counts=True)
        for v, c in zip(uniq, cnt):
            counts[format(int(v), f"0{prog0.n_clbits}b")] = int(c)
        if readout and "probabilities" in data:
            data["probabilities"] = DTCSimulator._readout_probs(data["probabilities"], readout)
        if readout and "expval_z_by_clbit" in data:
            for c, m in readout.items():
                if c in data["expval_z_by_clbit"]:
                    p0 = 0.5 * (1.0 + data["expval_z_by_clbit"][c])
                    data["expval_z_by_clbit"][c] = (p0 * m[0][0] + (1 - p0) * m[1][0]) - (p0 * m[0][1] + (1 - p0) * m[1][1])
        if "probabilities" in data:
            pr = data["probabilities"]
            data["expval_z"] = [sum(p * (1 - 2 * ((v >> c) & 1)) for v, p in pr.items())
                                for c in range(prog0.n_clbits)]
        elif "expval_z_by_clbit" in data:                 # unmeasured clbits read 0: <Z> = +1
            by = data.pop("expval_z_by_clbit")
            data["expval_z"] = [by.get(c, 1.0) for c in range(prog0.n_clbits)]
        data["counts"] = counts
        data["time_taken"] = time.time() - t0
        return ExperimentResult(name, counts, data, shots, seed)

    def sample_trajectories(self, circuit, traj_begin, traj_end, seed, noise_model=None):
        """Classical-register values (numpy int64) of noise trajectories traj_begin..traj_end-1 of `circuit`
        (one shot each).  The Philox counters are global trajectory ids, so disjoint ranges computed on
        different GPUs concatenate to exactly what one GPU produces (used by dist.ShardedSampler)."""
        torch = self.ctx.torch
        ctx = self.ctx
        nm = as_noise_model(self.noise_model if noise_model is None else noise_model)
        full_nm = nm
        if nm is not None and not nm.has_gate_noise():
            nm = None
        prog = self._compiled(as_circuit(circuit), nm)            # cached per (op list, noise model), like run()
        readout = {}
        if full_nm is not None and full_nm.has_readout_noise():
            readout = {int(c): m for c, m in ((c, full_nm.lookup_readout(q)) for c, q in prog.measure_orig.items()) if m is not None}
        meas = prog.measures
        k = len(meas)
        if k == 0 or k > MAX_PROB_QUBITS:
            raise ValueError("sample_trajectories needs 1..12 measured qubits")
        cbits = np.array([c for _, c in meas], dtype=np.int64)
        handle = capi.ProgramHandle(prog, ctx.index, self.engine)
        per = 16 << prog.n_main
        n = traj_end - traj_begin
        bt = max(1, min(n, (self.max_memory_bytes or int(0.7 * ctx.free_bytes())) // per))
        state = self._state_buffer(bt << prog.n_main)              # one buffer per simulator, reused by every call
        vals = np.zeros(n, dtype=np.int64)
        for a in range(0, n, bt):
            nt = min(bt, n - a)
            batch = evolve(ctx, prog, nt, traj_begin + a, seed, handle=handle, state=state, fused_rdm=True)
            cols = sample_rows(ctx, batch.outcome_probs(), 1, seed, traj_begin + a).cpu().numpy()[:, 0].astype(np.int64)
            v = np.zeros(nt, dtype=np.int64)
            for i in range(k):
                v |= ((cols >> i) & 1) << cbits[i]
            vals[a:a + nt] = self._readout_flips(v, readout, seed, np.arange(traj_begin + a, traj_begin + a + nt))
        handle.close()
        return vals

    @staticmethod
    def _clbit_probs(p_cols, to_clbits, n_clbits):
        vals = to_clbits(np.arange(len(p_cols)))
        out = {}
        for v, p in zip(vals, p_cols):
            out[int(v)] = out.get(int(v), 0.0) + float(p)
        return out


TARGET_OPERATIONS = ("cx", "id", "rz", "sx", "u1", "u2", "u3", "measure")


def make_backendv2_class(base=DTCSimulator):
    """DTCSimulator as a qiskit ``BackendV2``: the reference hands its backend to the transpiler --
    ``generate_preset_pass_manager(optimization_level=0, backend=backend, initial_layout=..., routing_method=None)``
    (fast.py:181-189) -- which needs ``backend.target``.  The Target lists exactly what AerSimulator offers once the
    reference's noise model is attached (NoiseModel basis gates id/rz/sx/cx + the noisy u1/u2/u3, plus measure) as
    ideal, all-to-all instructions on `num_qubits` >= 31 qubits (the scripts' snake layout uses physical index 30,
    fast.py:177; routing_method=None), so the level-0 lowering comes out as the committed gate_counts_*.csv show.
    Imports qiskit at call time; raises ImportError when it is absent."""
    from qiskit.circuit import Measure, Parameter
    from qiskit.circuit.library import CXGate, IGate, RZGate, SXGate, U1Gate, U2Gate, U3Gate
    from qiskit.providers import BackendV2, Options
    from qiskit.transpiler import Target

    class DTCSimulatorV2(base, BackendV2):
        __doc__ = base.__doc__

        def __init__(self, *args, num_qubits=None, **kw):
            BackendV2.__init__(self, provider=None, name=base.name, description="dtcsim B200-native DTC Floquet simulator",
                               backend_version="0.2.0")
            base.__init__(self, *args, **kw)
            self._n_target = int(num_qubits or base.num_qubits)
            self._target = None

        @property
        def target(self):
            if self._target is None:
                th, ph, lam = Parameter("theta"), Parameter("phi"), Parameter("lam")
                t = Target(num_qubits=self._n_target, description="dtcsim: ideal all-to-all basis of the reference's Aer setup")
                for gate in (CXGate(), IGate(), RZGate(lam), SXGate(), U1Gate(lam), U2Gate(ph, lam), U3Gate(th, ph, lam),
                             Measure()):
                    t.add_instruction(gate, None)          # properties None: available on every qubit (pair)
                self._target = t
            return self._target

        @property
        def max_circuits(self):
            return None

        @classmethod
        def _default_options(cls):
            return Options(shots=1024, seed_simulator=None, method="automatic")

        def run(self, run_input, **options):
            return base.run(self, run_input, **options)

        def set_options(self, **fields):
            return base.set_options(self, **fields)

    DTCSimulatorV2.__name__ = DTCSimulatorV2.__qualname__ = "DTCSimulator"
    return DTCSimulatorV2


DTCSimulatorBase = DTCSimulator
try:                                   # qiskit present: be a real BackendV2 (usable as backend= in the pass manager)
    DTCSimulator = make_backendv2_class(DTCSimulatorBase)
except ImportError:
    pass
AerSimulator = DTCSimulator
