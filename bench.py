#!/usr/bin/env python
"""bench.py -- Floquet periods/s on BASELINE.json config C2 (L=20 noisy forward+echo autocorrelation sweep,
t = 0..29, 1024 Pauli trajectories per circuit, complex128) with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = the whole sweep: 60 independent circuits (fast.py:217-257) x `--trajectories` trajectories
= 1305 x 1024 period applications.  `value` is device-resident throughput (programs compiled and
uploaded, state buffers allocated; timed with CUDA events); `e2e` runs the same sweep through the
public AerSimulator-compatible run() with host circuits in and counts out.  N > 1 (torchrun): weak
scaling over disorder instances (config C4) -- rank r simulates row r of hs_L20/phis_L20 -- with one
NCCL all-reduce of the autocorrelation sums at the end.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L, G, P_NOISE, QUBIT = 20, 0.97, 0.05, 10
METRIC = "floquet_periods_per_s"


# ------------------------------------------------------------------------------------------ workload
def load_disorder(row):
    import pandas as pd
    g = os.path.join(ROOT, "tests", "golden")
    hs = pd.read_csv(os.path.join(g, "hs_L20.csv")).values
    phis = pd.read_csv(os.path.join(g, "phis_L20.csv")).values
    return hs[row % len(hs)], phis[row % len(phis)]


def qc_circuit(dtcsim, hs, phis, t, echo):
    """The reference's qc_qiskit circuit body (fast.py:125-147) + its level-0 transpile (fast.py:176-190)."""
    circ = dtcsim.QuantumCircuit(L + 1, 1)
    circ.h(0)
    circ.cz(QUBIT + 1, 0)
    uf = dtcsim.QuantumCircuit(L + 1)
    for i in range(L):
        uf.rx(np.pi * G, i + 1)
    for i in range(0, L - 1, 2):
        uf.rzz(phis[i], i + 1, i + 2)
    for i in range(1, L - 1, 2):
        uf.rzz(phis[i], i + 1, i + 2)
    for i in range(L):
        uf.rz(hs[i], i + 1)
    for _ in range(t):
        circ.append(uf, range(L + 1))
    if echo:
        inv = uf.inverse()
        for _ in range(t):
            circ.append(inv, range(L + 1))
    circ.cz(QUBIT + 1, 0)
    circ.h(0)
    circ.measure(0, 0)
    pm = dtcsim.generate_preset_pass_manager(optimization_level=0, initial_layout=dtcsim.SNAKE_LAYOUT[:L + 1],
                                             routing_method=None)
    return pm.run(circ)


def sweep_points(tmax):
    return [(t, echo) for echo in (False, True) for t in range(tmax)]


def periods_of(points):
    return sum(t * (2 if echo else 1) for t, echo in points)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w": float(np.median(power)) if power else None}


# ------------------------------------------------------------------------------------------ reference arm / cpu baseline
def cpu_sample(hs, phis, sample_points, seed=1234):
    """The reference's CPU path restated (oracle/dtc_oracle.c): gate by gate, one trajectory per shot,
    complex128, OpenMP over amplitudes.  Returns (periods, seconds, cores)."""
    from oracle import c_oracle as CO
    from oracle import dtc_circuits as C
    from oracle import oracle as O
    CO.set_threads()                                  # all host cores, whatever OMP_NUM_THREADS the launcher exported
    noise = O.PauliNoise.depolarizing(P_NOISE)
    buf = np.empty(1 << (L + 1), dtype=np.complex128)
    jobs = []
    for t, echo in sample_points:
        ops, _, _ = C.autocorr_gates("vacuum", L, G, hs, phis, t, QUBIT, echo)
        oc, na, _ = O.compact_ops(C.lower_level0(ops, C.SNAKE_LAYOUT), 31)
        jobs.append((oc, na))
    t0 = time.perf_counter()
    for i, (oc, na) in enumerate(jobs):
        for tr in range(CPU_TRAJ):
            CO.run_trajectory(oc, na, noise, seed, i * CPU_TRAJ + tr, out=buf)
    dt = time.perf_counter() - t0
    return periods_of(sample_points) * CPU_TRAJ, dt, CO.threads()


CPU_SAMPLE = [(5, False), (5, True), (15, False), (15, True), (25, False)]   # 85 periods per trajectory
CPU_TRAJ = 6                                                                   # -> 510 periods, ~10 s on 16 host cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hs, phis = load_disorder(0)
    for _ in range(args.warmup):
        cpu_sample(hs, phis, CPU_SAMPLE[:1])
    tot_p, tot_s, cores = 0, 0.0, 1
    for _ in range(args.steps):
        p, s, cores = cpu_sample(hs, phis, CPU_SAMPLE)
        tot_p += p
        tot_s += s
    val = tot_p / tot_s
    sample = f"{len(CPU_SAMPLE)} circuits (t,echo)={CPU_SAMPLE}, {CPU_TRAJ} trajectories each, n=21, gate-by-gate"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "periods/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": "periods/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "periods/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"C2: L=20 (n=21) kicked-Ising DTC, g=0.97, depolarizing p=0.05 on u1/u2/u3, forward+echo "
                        f"autocorr t=0..{args.tmax - 1} ({2 * args.tmax} circuits), {args.trajectories} Pauli trajectories each",
            "tmax": args.tmax, "trajectories": args.trajectories, "state_bytes_n21": 16 << (L + 1),
            "register": "ancilla factorised out of the device register (n = 20, 16 MiB per trajectory); see DESIGN.md",
            "l2_policy": f"inputs larger than L2: every launch streams the whole batch ({args.trajectories} trajectories x "
                         f"{(16 << L) >> 20} MiB states of the n = {L} register = {(args.trajectories * (16 << L)) >> 30} GiB) once",
            "parallelism": f"disorder instances x{args.gpus} (weak), allreduce of sums"}



# ------------------------------------------------------------------------------------------ extra legs (sub-records)
def strong_leg(args, ctx, dist, rank, world, progs, handles, state, bt, points):
    """Strong scaling on the SAME job: config C2 on ONE disorder instance, the 1024 trajectories of every circuit split
    over the ranks (dist.shard_range; global Philox trajectory ids, so the counts equal the 1-GPU run), one all-reduce.
    Launches shrink by the number of ranks, so k_frames, the read-out kernels and the host enqueue weigh more."""
    import torch
    from dtcsim import backend, dist as D
    NT = args.trajectories
    a0, b0 = D.shard_range(NT, rank, world)
    sums = torch.zeros(len(points), 2, dtype=torch.float64, device=ctx.device)

    def step(seed):
        sums.zero_()
        pending = []
        for i, (prog, h) in enumerate(zip(progs, handles)):
            for a in range(a0, b0, bt):
                nt = min(bt, b0 - a)
                batch = backend.evolve(ctx, prog, nt, a, seed + i, handle=h, state=state, fused_rdm=True)
                pending.append((i, batch, batch.readout_rdm() if prog.small else None))
        for i, batch, rdm in pending:
            pr = batch.outcome_probs(rdm)
            ez = pr[:, 0] - pr[:, 1]
            sums[i, 0] += ez.sum()
            sums[i, 1] += (ez * ez).sum()
        dist.all_reduce(sums)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    step(900)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    step(1234)
    enq = time.perf_counter() - t0                      # host time to enqueue the whole step
    e1.record()
    sync_all()
    sweep_ms = 0.0
    for h, prog in zip(handles, progs):
        if prog.n_main >= 12:
            sweep_ms += h.pass_time()[0]
    ms = torch.tensor([e0.elapsed_time(e1), sweep_ms, enq * 1e3], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step, sweep_ms, enq_ms = (float(x) for x in ms.tolist())
    periods = periods_of(points) * NT
    return {"scaling": "strong", "workload": "C2 on one disorder instance, trajectories of each circuit split over the ranks",
            "value": periods / (ms_step * 1e-3), "unit": "periods/s", "ms_per_step": ms_step, "steps": 1,
            "trajectories_per_rank": b0 - a0,
            "sweep_kernels_ms": sweep_ms, "non_sweep_share": max(0.0, 1.0 - sweep_ms / ms_step),
            "host_enqueue_ms": enq_ms,
            "note": "non_sweep_share = 1 - (k_tile_stream time / step time): k_frames, read-out kernels, launch gaps, all-reduce"}


def sharded_circuit(dtcsim, L, g, hs, phis, t, echo):
    """dtc_qasm.py:70-91 circuit shape (L qubits, no ancilla): t periods, optionally followed by their inverse."""
    c = dtcsim.QuantumCircuit(L, L)
    uf = dtcsim.QuantumCircuit(L)
    for i in range(L):
        uf.rx(np.pi * g, i)
    for i in range(0, L - 1, 2):
        uf.rzz(phis[i], i, i + 1)
    for i in range(1, L - 1, 2):
        uf.rzz(phis[i], i, i + 1)
    for i in range(L):
        uf.rz(hs[i], i)
    for _ in range(t):
        c.append(uf, range(L))
    if echo:
        inv = uf.inverse()
        for _ in range(t):
            c.append(inv, range(L))
    c.measure_all()
    return dtcsim.lower_level0(c)


def sharded_leg(args, dist, rank, world, local, hbm_peak):
    """Config C5: one statevector of L = 31 + log2(P) qubits sharded on its top log2(P) qubits (32 GiB per GPU), 10 periods
    forward + 10 inverse (so every <Z_q> must return to +1 exactly), and an L = 22 noisy trajectory against the C oracle."""
    import torch
    import dtcsim
    from dtcsim import sharded
    g = int(round(np.log2(world)))
    out = {}

    def allred_on(dev):
        def f(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            dist.all_reduce(t)
            return t.cpu().numpy()
        return f

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    # ---- parity at a size the oracle finishes in seconds: L = 22, one noisy trajectory, vs the gate-by-gate C oracle
    Lc = 22
    rng = np.random.default_rng(Lc)
    hs = rng.random(Lc) * 2 * np.pi - np.pi                      # generate_disorder.py:16-18
    phis = rng.random(Lc - 1) * np.pi - 1.5 * np.pi
    circ = sharded_circuit(dtcsim, Lc, 0.97, hs, phis, 3, False)
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(dtcsim.depolarizing_error(0.05, 1), ["u1", "u2", "u3"], warnings=False)
    eng = sharded.CudaShardEngine(Lc, Lc - g, rank, world, local)
    res = sharded.ShardedStatevector(Lc, rank, world, eng, all_reduce=allred_on(eng.ctx.device)).run(circ, nm, seed=1234, trajectory=3)
    transport = eng.transport
    eng.close()
    del eng
    dev = torch.zeros(1, dtype=torch.float64, device=torch.device("cuda", local))
    if rank == 0:
        from oracle import c_oracle as CO
        from oracle import oracle as O
        oc, na, _ = O.compact_ops([o.astuple() for o in circ.ops], Lc)
        psi = CO.run_trajectory(oc, na, O.PauliNoise.depolarizing(0.05), 1234, 3)
        p = np.abs(psi) ** 2
        idx = np.arange(1 << Lc)
        want = np.array([np.sum(p * (1.0 - 2.0 * ((idx >> q) & 1))) for q in range(Lc)])
        dev[0] = float(np.abs(want - np.array(res["expect_z"])).max())
    dist.all_reduce(dev)
    out["oracle_max_dev"] = float(dev.item())
    out["oracle_check"] = f"L={Lc} noisy trajectory (p=0.05, 3 periods), <Z_q> of all qubits vs the gate-by-gate C oracle"

    # ---- the timed run
    L = 31 + g
    nl = L - g
    rng = np.random.default_rng(34)
    hs = rng.random(L) * 2 * np.pi - np.pi
    phis = rng.random(L - 1) * np.pi - 1.5 * np.pi
    torch.cuda.empty_cache()
    eng = sharded.CudaShardEngine(L, nl, rank, world, local)
    allred = allred_on(eng.ctx.device)
    sharded.ShardedStatevector(L, rank, world, eng, all_reduce=allred).run(sharded_circuit(dtcsim, L, 0.97, hs, phis, 1, True))
    T = args.sharded_periods
    circ = sharded_circuit(dtcsim, L, 0.97, hs, phis, T, True)
    sv = sharded.ShardedStatevector(L, rank, world, eng, all_reduce=allred)
    eng.passes = 0
    eng.passes_weighted = 0.0
    eng.sliced_exchanges = 0
    sync_all()
    t0 = time.perf_counter()
    res = sv.run(circ)
    sync_all()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.ctx.device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    ez = np.array(res["expect_z"])
    periods = 2 * T
    shard_bytes = 16 << nl
    wire = shard_bytes * (world - 1) // world
    t_hbm = 2 * shard_bytes / (hbm_peak * 1e9)               # one read + one write of the shard per period
    t_link = wire / 770e9                                    # measured NVLink peer-copy bandwidth per direction
    out.update({"workload": f"C5: L={L} statevector sharded on its top {g} qubits ({shard_bytes >> 30} GiB per GPU), "
                            f"{T} periods forward + {T} inverse, complex128",
                "value": periods / dt, "unit": "periods/s", "seconds": dt, "ms_per_period": 1e3 * dt / periods,
                "norm": res["norm"], "echo_max_dev": float(np.abs(ez - 1).max()),
                "state_sweeps_per_period": eng.passes_weighted / periods, "exchanges": sv.stats["exchanges"],
                "exchanges_overlapped_with_sweeps": eng.sliced_exchanges, "transport": transport,
                "exchange_gbs_per_rank_per_direction": sv.stats["exchange_bytes_per_rank"] / dt / 1e9,
                "hbm_algorithmic_gbs_per_rank": eng.passes_weighted * 2 * shard_bytes / dt / 1e9,
                "floor_ms_per_period": {"hbm": 1e3 * t_hbm, "nvlink_770": 1e3 * t_link},
                "frac_of_floor": max(t_hbm, t_link) / (dt / periods),
                "timer": "wall clock between barriers, max over ranks (includes the final <Z> reduction)"})
    out["ok"] = bool(abs(res["norm"] - 1) < 1e-9 and np.abs(ez - 1).max() < 1e-9 and out["oracle_max_dev"] < 1e-10)
    eng.close()
    del eng, sv
    torch.cuda.empty_cache()
    return out


def sweep_leg(args, ctx, dist, rank, world, noise):
    """Config C4 through the product front-end dtcsim.run_sweep: g values (generate_params.py:6) x polarisations
    (pol.py:336) x t x echo on disorder row 0, points dealt over the ranks, one all-reduce."""
    import torch
    import dtcsim
    from dtcsim import sweeps
    g_list = list(sweeps.G_GRID) if args.sweep_g == "grid" else [float(x) for x in args.sweep_g.split(",")]
    pols = args.sweep_pol.split(",")
    hs, phis = load_disorder(0)
    sim = dtcsim.AerSimulator(noise_model=noise, device="GPU", cuStateVec_enable=True, cuda_device=ctx.index)
    t_values = list(range(args.sweep_tmax))
    rec = {}
    for pol in pols:                                  # one call per polarisation so that passes / period can be reported
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        r = dtcsim.run_sweep(sim, L, g_list, hs[None, :], phis[None, :], t_values, (False, True), (pol,), shots=args.trajectories,
                             seed_simulator=1234, rank=rank, world=world)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        dt = time.perf_counter() - t0
        c = sweeps.autocorr_circuit(L, g_list[0], hs, phis, 4, echo=False, polarization=pol)
        prog = dtcsim.compile_circuit(c, dtcsim.as_noise_model(noise), optimize=True)
        from dtcsim import capi
        h = capi.ProgramHandle(prog, ctx.index)
        rec[pol] = {"periods_per_s": r["periods"] / dt, "seconds": dt, "points": r["points"],
                    "sweeps_per_period_t4": h.num_passes / 4.0,
                    "autocorr_g0_t1": float(r["mean"][0, 0, 0, 1]) if len(t_values) > 1 else None}
        h.close()
    return {"workload": f"C4: L=20, g in {g_list}, pol in {pols}, t=0..{args.sweep_tmax - 1} forward+echo, "
                        f"{args.trajectories} trajectories per point, points dealt over {world} rank(s)",
            "by_polarization": rec}


# ------------------------------------------------------------------------------------------ config C3 (exact density matrix)
def c3_record(args, device=0):
    """BASELINE config C3: L = 12 exact noisy density-matrix evolution (rho = 2^24 complex128 = 256 MiB), 20 periods,
    hs_L20/phis_L20 row 0 entries 0..11, g = 0.97, p = 0.05, observable <Z_6>.  One step = the 20-period evolution.
    Returns the JSON record (`bench.py --config C3` prints it; the default C2 run carries it as the `c3` sub-record)."""
    import torch
    import dtcsim
    from dtcsim import backend
    n, T = 12, 20
    hs, phis = load_disorder(0)
    hs, phis = hs[:n], phis[:n - 1]
    c = dtcsim.QuantumCircuit(n, 1)
    for _ in range(T):
        for i in range(n):
            c.rx(np.pi * G, i)
        for i in range(0, n - 1, 2):
            c.rzz(phis[i], i, i + 1)
        for i in range(1, n - 1, 2):
            c.rzz(phis[i], i, i + 1)
        for i in range(n):
            c.rz(hs[i], i)
    c.measure(6, 0)
    noise = dtcsim.NoiseModel()
    noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(P_NOISE, 1), ["u1", "u2", "u3"], warnings=False)
    prog = dtcsim.compile_circuit(dtcsim.lower_level0(c), dtcsim.as_noise_model(noise), want_dm=True)
    ctx = backend.DeviceContext(device)
    stats = {}
    n_warm = max(args.warmup, 20)                        # ~0.15 s: a cold GPU needs that long to reach its clocks
    for _ in range(n_warm):
        rho = backend.run_density_matrix(ctx, prog, stats)
    torch.cuda.synchronize()
    sampler = ClockSampler(device)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(args.steps, 50)
    e0.record()
    for _ in range(steps):
        rho = backend.run_density_matrix(ctx, prog, stats)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    d = 1 << n
    diag = rho.view(d, d).diagonal().real.cpu().numpy()
    ez = float(np.sum(diag * (1.0 - 2.0 * ((np.arange(d) >> prog.bit_of[6]) & 1))))
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_file))["hbm_gbs"]) if os.path.exists(peaks_file) else 6650.0
    b_alg = 2 * 16 * (1 << (2 * n))                       # one read + one write of rho per period
    value = T / (ms * 1e-3)
    ach = value * b_alg / 1e9
    sweeps = stats["sweeps"]
    line = {"metric": METRIC, "value": value, "unit": "periods/s", "n_gpus": 1, "steps": steps, "warmup": n_warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic",
            "config": {"workload": "C3: L=12 exact noisy density matrix (2^24 complex128 = 256 MiB), g=0.97, depolarizing p=0.05, "
                                   "20 periods, <Z_6>", "l2_policy": "rho (256 MiB) is larger than L2"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "kernel": "k_dm_reg", "sweeps_per_period": sweeps / T,
                         "per_sweep_gbs": sweeps * b_alg / (ms * 1e-3) / 1e9,
                         "note": "achieved = periods/s x 2 x 16 B x 4^12 (SURVEY 8d byte model: ONE read + write of rho per period); "
                                 "per_sweep_gbs = what each of the sweeps_per_period passes over rho sustains"},
            "expect_z6_after_20_periods": ez, "trace": float(diag.sum()), "gpu_launches": steps * (2 * sweeps + 2), "clocks": clocks}
    return line


def run_c3(args):
    print(json.dumps(c3_record(args)), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import dtcsim
    from dtcsim import backend, capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = backend.DeviceContext(local)
    capi.RESIDENT = bool(args.resident)
    if args.resident_mb:
        capi.set_resident_bytes(args.resident_mb << 20)
    hs, phis = load_disorder(rank)                      # weak scaling: one disorder instance per rank (C4)
    points = sweep_points(args.tmax)
    noise = dtcsim.NoiseModel()
    noise.add_all_qubit_quantum_error(dtcsim.depolarizing_error(P_NOISE, 1), ["u1", "u2", "u3"], warnings=False)
    circuits = [qc_circuit(dtcsim, hs, phis, t, echo) for t, echo in points]
    nm = dtcsim.as_noise_model(noise)
    progs = [dtcsim.compile_circuit(c, nm, optimize=True) for c in circuits]   # ancilla factorised: register n = 20
    handles = [capi.ProgramHandle(p, ctx.index) for p in progs]
    for h in handles:
        h.set_profiling(True)
    NT = args.trajectories
    nmax = max(p.n_main for p in progs)
    per = 16 << nmax
    bt = max(1, min(NT, int(0.6 * ctx.free_bytes()) // per))
    state = ctx.empty(bt << nmax, torch.complex128)
    sums = torch.zeros(len(points), 2, dtype=torch.float64, device=ctx.device)
    launches = [0]
    resident = [0]

    def step(seed):
        sums.zero_()
        pending = []
        for i, (prog, h) in enumerate(zip(progs, handles)):
            for a in range(0, NT, bt):
                nt = min(bt, NT - a)
                batch = backend.evolve(ctx, prog, nt, a, seed + i, handle=h, state=state, fused_rdm=True)
                # read-out: the density matrix of site q comes out of the last pass (fused) or of a dtc_rdm reduction
                # taken now (the state buffer is reused by the next circuit); k_readout_small finishes it below
                pending.append((i, batch, batch.readout_rdm() if prog.small else None))
                # kernels of this run (frames + sweeps: one persistent launch in resident execution) + read-out kernels
                launches[0] += h.last_run_info()[1] + (1 if batch.fused_rdm is not None else 2)
                resident[0] += int(getattr(batch, "resident", False))
        for i, batch, rdm in pending:
            pr = batch.outcome_probs(rdm)
            ez = pr[:, 0] - pr[:, 1]
            sums[i, 0] += ez.sum()
            sums[i, 1] += (ez * ez).sum()
        if dist is not None:
            dist.all_reduce(sums)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for w in range(args.warmup):
        step(1000 + w)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches[0] = 0
    resident[0] = 0
    pass_ms, pass_n, half_passes, n_timed_launches = 0.0, 0, 0, 0
    by_mode = {}
    e0.record()
    for k in range(args.steps):
        step(1234 + k)
        if k == args.steps - 1:
            e1.record()
    sync_all()
    for h, prog in zip(handles, progs):                 # kernel-only time of the last step's sweeps
        if prog.n_main >= 12:
            ms, n = h.pass_time()
            pass_ms += ms
            pass_n += n
            n_timed_launches += 1 if h.last_run_info()[0] else n
            if not h.last_run_info()[0]:
                pt = h.pass_times()
                for j, (pms, mode) in enumerate(pt):
                    if 0 < j < len(pt) - 1:                 # full-traffic sweeps only (not the generated first / fused last one)
                        by_mode.setdefault(mode, [0.0, 0])
                        by_mode[mode][0] += pms
                        by_mode[mode][1] += 1
            half_passes += sum(h.last_run_flags())        # passes that only write (generated start) or only read (fused read-out)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=ctx.device)
    if dist is not None:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    periods_step = periods_of(points) * NT * world
    value = periods_step / (ms_step * 1e-3)
    gpu_launches = launches[0]
    autocorr = (sums[:, 0] / (NT * world)).cpu().numpy()

    # ---- roofline of the dominant kernel (k_tile_stream, the TMA-fed fused pass; k_tile_pass where a pass is not
    #      eligible): algorithmic bytes per launch / launch duration
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        with open(peaks_file) as fh:
            peak, peak_src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    last_batch = NT - ((NT - 1) // bt) * bt              # pass_time() refers to each handle's last run
    bytes_per_launch = 2 * 16 * (1 << nmax) * last_batch           # one read + one write of the batch of states
    roof = None
    n_stream = int(sum(h.num_stream_passes for h in handles))
    n_pass_all = int(sum(h.num_passes for h in handles))
    traffic, traffic_src = None, None
    tf = os.path.join(ROOT, "profiles", "ncu_tile_resident_traffic.json" if resident[0] > 0 else "ncu_tile_stream_traffic.json")
    if os.path.exists(tf):
        with open(tf) as fh:
            tj = json.load(fh)
        traffic = float(tj["dram_bytes_per_algorithmic_byte"]) * bytes_per_launch   # per full-traffic sweep
        traffic_src = ("not measured in this run: DRAM bytes / algorithmic byte of the committed ncu --set full capture "
                       f"({tj.get('source', 'profiles/')}) x this run's bytes_per_launch")
    is_resident = resident[0] > 0
    if pass_n:
        avg_ms = pass_ms / max(n_timed_launches, 1)
        # algorithmic bytes of the timed launches: a full pass reads and writes the batch once; the first pass of a
        # circuit only writes it (generated start) and a fused last pass only reads it
        alg_bytes = bytes_per_launch * (pass_n - 0.5 * half_passes)
        ach = alg_bytes / (pass_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "kernel": "k_tile_resident" if is_resident else ("k_tile_stream" if 2 * n_stream >= n_pass_all else "k_tile_pass"),
                "stream_passes": n_stream, "passes": n_pass_all, "avg_launch_ms": avg_ms, "launches_timed": n_timed_launches,
                "sweeps_timed": pass_n, "bytes_per_sweep": bytes_per_launch,
                "bytes_per_launch": alg_bytes / max(n_timed_launches, 1), "half_traffic_sweeps": half_passes,
                "execution": ("resident: one persistent launch per circuit runs all its sweeps over groups of trajectories whose "
                              "states stay in L2 (algorithmic bytes are what a sweep-per-pass execution moves through HBM; see "
                              "traffic for the DRAM bytes actually moved)") if is_resident else "one launch per sweep, batch streamed through HBM",
                "algorithmic_bytes_timed": alg_bytes, "peak_source": peak_src,
                "register_qubits": nmax, "traffic_source": traffic_src,
                "by_tile_layout": {{1: "contiguous_64KB_tiles", 2: "runs_of_64B", 3: "runs_of_2KB", 0: "register_fed"}.get(m, str(m)):
                                   {"launches": v[1], "avg_launch_ms": v[0] / v[1],
                                    "gbs": bytes_per_launch / (v[0] / v[1] * 1e-3) / 1e9, "frac": bytes_per_launch / (v[0] / v[1] * 1e-3) / 1e9 / peak}
                                   for m, v in sorted(by_mode.items()) if v[1]},
                "periods_frac_actual_register": (value / world) * (2 * 16 * (1 << nmax)) / (peak * 1e9)}

    # ---- end to end through the public API (host circuits in, counts out)
    e2e = None
    if not args.no_e2e:
        sim = dtcsim.AerSimulator(noise_model=noise, device="GPU", cuStateVec_enable=True, cuda_device=local)
        h2d = sum(sum(a.nbytes for a in p.arrays().values()) + p.n_layers * 4616 for p in progs)
        d2h = len(points) * NT * (4 + 16)

        def e2e_pass(seed):
            # one run() call on the list of 60 host circuits (Aer accepts a list); the backend pipelines them
            res = sim.run(circuits, shots=NT, seed_simulator=seed).result()
            return [backend.compute_z_expectation(res.get_counts(c), 1)[0] for c in circuits]

        e2e_pass(77)
        sync_all()
        # the timed pass pays the host compile of every circuit, as a real sweep does (it never repeats a circuit):
        # drop the programs the warm-up pass left in the simulator's cache
        sim.__dict__.get("_prog_cache", {}).clear()
        t0 = time.perf_counter()
        res = e2e_pass(1234)
        sync_all()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=ctx.device)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": periods_step / float(dt.item()), "unit": "periods/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "seconds": float(dt.item()), "timer": "host wall clock around one run(list of 60 host circuits) call + get_counts; program cache cleared first, "
                        "so the host compile, table upload, frames, sweeps, read-out and the device->host copies are all inside",
               "autocorr_t1": res[1] if len(res) > 1 else None}

    extra = {}
    if world > 1 and not args.no_strong:
        extra["strong"] = strong_leg(args, ctx, dist, rank, world, progs, handles, state, bt, points)
    if args.sweep_g:
        extra["c4_sweep"] = sweep_leg(args, ctx, dist, rank, world, noise)
    if world > 1 and not args.no_sharded:
        # the C2 buffers go first: the sharded state takes 64 GiB per GPU
        for h in handles:
            h.close()
        del state, sums
        if not args.no_e2e:
            sim._state = None
        torch.cuda.empty_cache()
        try:
            extra["sharded"] = sharded_leg(args, dist, rank, world, local, peak)
        except Exception as exc:                      # the headline line must still be printed
            extra["sharded"] = {"error": repr(exc)[:400]}
    if world == 1 and not args.no_c3:
        # config C3 (exact density matrix, fits one GPU: replicas only) rides along as a sub-record of the 1-GPU run
        try:
            c3 = c3_record(args, local)
            extra["c3"] = {k: c3[k] for k in ("value", "unit", "steps", "ms_per_step", "config", "roofline",
                                              "expect_z6_after_20_periods", "trace", "gpu_launches")}
        except Exception as exc:
            extra["c3"] = {"error": repr(exc)[:400]}
    if rank == 0:
        cpu = None
        if not args.no_cpu:
            p, s, cores = cpu_sample(hs, phis, CPU_SAMPLE)
            cpu = {"value": p / s, "unit": "periods/s", "cores": cores, "kind": "port",
                   "sample": f"{len(CPU_SAMPLE)} circuits {CPU_SAMPLE}, {CPU_TRAJ} trajectories each ({p} periods), n=21, gate-by-gate C/OpenMP oracle"}
        line = {"metric": METRIC, "value": value, "unit": "periods/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "complex128", "data": "synthetic", "config": workload_config(args),
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
                "passes_per_sweep": n_pass_all, "periods_per_sweep": periods_of(points),
                "autocorr_forward_t1_t2": [float(autocorr[1]), float(autocorr[2])] if args.tmax > 2 else None}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trajectories", type=int, default=1024)
    ap.add_argument("--tmax", type=int, default=30, help="sweep t = 0..tmax-1 (30 = BASELINE config)")
    ap.add_argument("--config", default="C2", choices=["C2", "C3"], help="C2 (default, the headline) or C3 (exact density matrix)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--resident", action="store_true", help="resident execution: all sweeps of a circuit in one persistent launch over L2-resident trajectory groups")
    ap.add_argument("--resident-mb", type=int, default=0, help="state MiB kept in flight per group in resident execution (default 64)")
    ap.add_argument("--no-c3", action="store_true", help="skip the C3 (exact density matrix) sub-record")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling sub-record")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sharded-statevector (C5) sub-record")
    ap.add_argument("--sharded-periods", type=int, default=10, help="C5: periods forward (+ the same number inverse)")
    ap.add_argument("--sweep-g", default="", help="C4 sub-record: 'grid' (generate_params.py:6) or comma-separated g values")
    ap.add_argument("--sweep-pol", default="x,y,xy,yx", help="C4 sub-record: polarisations (pol.py:336)")
    ap.add_argument("--sweep-tmax", type=int, default=10, help="C4 sub-record: t = 0..tmax-1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "C3":
        run_c3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
