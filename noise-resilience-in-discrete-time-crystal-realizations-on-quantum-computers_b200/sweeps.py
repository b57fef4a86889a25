"""Batched sweep front-end: the loops the reference's drivers run one circuit at a time.

The reference evaluates a sweep point by point -- `get_single_out` / `get_instances` loop over t, echo and disorder
instances (fast.py:217-239), `for pol in ["x","y","xy","yx"]` over polarisations (pol.py:336,349-373), the parameter
grid over g (generate_params.py:6) -- and every point pays backend construction + transpile + one `run()`
(fast.py:156,181-192,211).  `run_sweep` takes the whole grid at once: it builds the same circuits (same gate sequence,
same level-0 lowering, same snake layout), deals the points to the ranks of the process group (dist.deal_units), runs
each rank's share through ONE pipelined `run(list)` call and sums the results with one all-reduce.  The numbers are
what the per-point loop returns: <Z_ancilla> from the counts of `shots` Pauli trajectories per point
(compute_z_expectation, fast.py:92-109), seeds `seed + point index`, independent of the number of ranks.
"""
import math

import numpy as np

from .ir import QuantumCircuit
from .lowering import SNAKE_LAYOUT, lower_level0

PI = math.pi
POLARIZATIONS = ("x", "y", "xy", "yx", "circular_left", "circular_right", "circular_static", "xy_cycle")
G_GRID = (0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95, 1.0)      # generate_params.py:6


def floquet_period(L, g, phis, hs, polarization="x", time_step=0, circular_frequency=1.0):
    """One period U_F on circuit qubits 1..L (qubit 0 is the ancilla): kick layer in the given polarisation
    (pol.py:110-122, circ-pol.py:110-140), RZZ on even then odd bonds, RZ fields (fast.py:111-121)."""
    sub = QuantumCircuit(L + 1)
    for i in range(L):
        q = i + 1
        if polarization == "x":
            sub.rx(PI * g, q)
        elif polarization == "y":
            sub.ry(PI * g, q)
        elif polarization == "xy":
            sub.rx(PI * g / 2, q)
            sub.ry(PI * g / 2, q)
        elif polarization == "yx":
            sub.ry(PI * g / 2, q)
            sub.rx(PI * g / 2, q)
        elif polarization in ("circular_left", "circular_right"):
            sgn = 1.0 if polarization == "circular_left" else -1.0
            sub.rx(PI * g * math.cos(circular_frequency * time_step) / math.sqrt(2), q)
            sub.ry(sgn * PI * g * math.sin(circular_frequency * time_step) / math.sqrt(2), q)
        elif polarization == "circular_static":
            sub.rx(PI * g / math.sqrt(2), q)
            sub.ry(PI * g / math.sqrt(2), q)
        else:
            raise ValueError(f"unknown polarization {polarization!r}")
    for i in range(0, L - 1, 2):
        sub.rzz(float(phis[i]), i + 1, i + 2)
    for i in range(1, L - 1, 2):
        sub.rzz(float(phis[i]), i + 1, i + 2)
    for i in range(L):
        sub.rz(float(hs[i]), i + 1)
    return sub


def autocorr_circuit(L, g, hs, phis, t, qubit=None, echo=False, polarization="x", initial_state="vacuum",
                     g_values=None, pol_schedule=None, circular_frequency=1.0, transpile=True):
    """The reference's Hadamard-test autocorrelation circuit (fast.py:124-147) for t periods.

    g_values: per-step g (ctrl-g.py:196-241; step k uses g_values[k], the echo undoes the steps in reverse order);
    pol_schedule: callable step -> polarisation (xy-cycle.py:141-157).  transpile=True applies the level-0 lowering
    with the snake layout (fast.py:176-190), i.e. returns what the scripts hand to backend.run()."""
    qubit = L // 2 if qubit is None else int(qubit)
    circ = QuantumCircuit(L + 1, 1)
    if initial_state == "neel":
        for i in range(1, L + 1):
            if i % 2 == 0:
                circ.x(i)
    elif initial_state != "vacuum":
        raise ValueError(f"unknown initial state {initial_state!r}")
    circ.h(0)
    circ.cz(qubit + 1, 0)

    if polarization == "xy_cycle" and pol_schedule is None:        # xy-cycle.py:144-156 as a named polarisation
        pol_schedule = xy_cycle_schedule()

    def period(step):
        gg = g if g_values is None else g_values[step]
        pol = polarization if pol_schedule is None else pol_schedule(step)
        return floquet_period(L, gg, phis, hs, pol, step, circular_frequency)

    for step in range(t):
        circ.append(period(step), range(L + 1))
    if echo:
        for step in range(t - 1, -1, -1):
            circ.append(period(step).inverse(), range(L + 1))
    circ.cz(qubit + 1, 0)
    circ.h(0)
    circ.measure(0, 0)
    if not transpile:
        return circ
    if L + 1 > len(SNAKE_LAYOUT):
        return lower_level0(circ)
    return lower_level0(circ, SNAKE_LAYOUT[:L + 1])


def xy_cycle_schedule(block=5):
    """Polarisation alternating x / y every `block` steps (xy-cycle.py:144-156)."""
    return lambda step: "x" if (step // block) % 2 == 0 else "y"


def sweep_points(g_list, polarizations, instances, t_values, echoes):
    """Flattened grid in the order the reference's nested loops visit it: g, polarisation, echo, instance, t."""
    return [(gi, pi, ei, ii, ti) for gi in range(len(g_list)) for pi in range(len(polarizations))
            for ei in range(len(echoes)) for ii in range(len(instances)) for ti in range(len(t_values))]


def run_sweep(sim, L, g_list, hs, phis, t_values, echoes=(False, True), polarizations=("x",), qubit=None,
              initial_state="vacuum", shots=1024, seed_simulator=1234, rank=0, world=1, group=None,
              circular_frequency=1.0, chunk=64):
    """Whole autocorrelation sweep in one call.

    sim: DTCSimulator (with its noise model) bound to this rank's GPU.  hs [n_inst, >= L], phis [n_inst, >= L-1]: disorder
    rows (hs_L*.csv / phis_L*.csv).  Returns a dict with
      "autocorr"  float64 [len(g_list), len(polarizations), len(echoes), n_inst, len(t_values)]: <Z_ancilla> per point
                  from counts, as `qc_qiskit` returns it (fast.py:213);
      "mean"      the instance average [g, pol, echo, t]  (get_instances, fast.py:228-239);
      "points", "periods": number of circuits and of Floquet periods x shots simulated over all ranks.
    With world > 1 every rank must call it (one all-reduce at the end); results are identical on all ranks and
    identical to a single-rank call with the same seed."""
    from . import dist as D
    from .backend import compute_z_expectation
    hs = np.atleast_2d(np.asarray(hs, dtype=np.float64))
    phis = np.atleast_2d(np.asarray(phis, dtype=np.float64))
    n_inst = hs.shape[0]
    pts = sweep_points(g_list, polarizations, range(n_inst), t_values, echoes)
    mine = D.deal_units(len(pts), rank, world)
    shape = (len(g_list), len(polarizations), len(echoes), n_inst, len(t_values))
    out = np.zeros(shape, dtype=np.float64)
    periods = 0
    for a in range(0, len(mine), chunk):                      # bounded host memory: circuits are built chunk by chunk
        idx = mine[a:a + chunk]
        circs = []
        for k in idx:
            gi, pi, ei, ii, ti = pts[k]
            t = int(t_values[ti])
            circs.append(autocorr_circuit(L, g_list[gi], hs[ii], phis[ii], t, qubit, echoes[ei], polarizations[pi],
                                          initial_state, circular_frequency=circular_frequency))
            periods += t * (2 if echoes[ei] else 1) * shots
        if not circs:
            continue
        # per-circuit seeds = seed + GLOBAL point index, so the result does not depend on how the points are dealt
        res = sim.run(circs, shots=shots, seed_simulator=[int(seed_simulator) + int(k) for k in idx]).result()
        for j, k in enumerate(idx):
            out[pts[k]] = compute_z_expectation(res.get_counts(j), 1)[0]
    tot = np.concatenate([out.reshape(-1), [float(periods)]])
    tot = D.all_reduce_sum(tot, group, sim.ctx.device if world > 1 else None)
    out = tot[:-1].reshape(shape)
    return {"autocorr": out, "mean": out.mean(axis=3), "points": len(pts), "periods": int(round(tot[-1]))}


# ----------------------------------------------------------------------------------- real-time adaptive g (ctrl-g.py, g-opt.py)
def feedback_g(echo_val, target_echo, current_g, time_step, feedback_gain, g_min, g_max, exponential=True,
               decay_compensation=0.1):
    """Next-step g from the echo just measured: `calculate_exponential_g_adjustment` (ctrl-g.py:352-398, g-opt.py:429-474).

    linear:       g + gain * (target - echo)
    exponential:  [gain * error * exp(c t) + log term] * (1 + c t) added to g, the log term being gain * 0.1 * ln(target / echo)
                  for 0.01 < echo < target, 0 for echo >= target and 2 * gain for echo <= 0.01;  clipped to [g_min, g_max]."""
    err = target_echo - echo_val
    if exponential:
        adj = feedback_gain * err * math.exp(decay_compensation * time_step)
        if echo_val > 0.01:
            adj += feedback_gain * (math.log(target_echo / echo_val) if echo_val < target_echo else 0.0) * 0.1
        else:
            adj += feedback_gain * 2.0
        new_g = current_g + adj * (1.0 + decay_compensation * time_step)
    else:
        new_g = current_g + feedback_gain * err
    return float(min(max(new_g, g_min), g_max))


def optimize_g(echo_of, target_echo, g_min, g_max, method="bounded", grid_points=10, xatol=1e-5, maxiter=500):
    """g in [g_min, g_max] minimising (echo(g) - target)^2: `optimize_g_for_target_echo` (g-opt.py:354-393).

    echo_of(list of g) -> list of echo values (one circuit per candidate; a list lets the grid go through ONE run(list)).
    method "bounded": Brent's bounded scalar minimiser as the reference calls it (scipy `minimize_scalar(method='bounded')`,
    one candidate per call -- sequential by construction); "grid": the reference's fallback `grid_search_g_optimization`
    (g-opt.py:395-428: `grid_points` equidistant candidates, smallest |echo - target| wins, first one on ties)."""
    if method == "bounded":
        from scipy.optimize import minimize_scalar
        res = minimize_scalar(lambda g: (echo_of([float(g)])[0] - target_echo) ** 2, bounds=(g_min, g_max), method="bounded",
                              options={"xatol": xatol, "maxiter": maxiter})
        if res.success:
            return float(res.x)
    elif method != "grid":
        raise ValueError(f"unknown optimiser {method!r}")
    cands = [float(x) for x in np.linspace(g_min, g_max, grid_points)]
    dist = [abs(e - target_echo) for e in echo_of(cands)]
    return cands[int(np.argmin(dist))]


def run_adaptive(sim, L, hs, phis, T, g_initial=0.84, target_echo=1.0, feedback_gain=0.01, exponential_feedback=True,
                 decay_compensation=0.1, g_min=0.84, g_max=1.0, use_optimization=False, optimizer="bounded", grid_points=10,
                 qubit=None, initial_state="vacuum", shots=1024, seed_simulator=1234, rank=0, world=1, group=None,
                 parallel="instances"):
    """Real-time adaptive control of the kick strength: `get_instances_adaptive_realtime` (ctrl-g.py:443-490; with
    use_optimization g-opt.py:500-556).

    Per disorder instance, for t = 0..T-1: the circuits of step t use g_history + [g_t] (time-dependent g, ctrl-g.py:196-241)
    for t+1 periods; forward and echo are evaluated (one pipelined run([forward, echo])), and g_{t+1} follows from the echo by
    `feedback_g`, or -- use_optimization -- from `optimize_g` over the g of step t (the reference hands the optimiser the
    history WITHOUT step t and t+1 periods, and uses the optimum as the next step's g; kept as is).

    The steps of one instance depend on each other; what runs in parallel is (a) forward / echo / grid candidates of a step in
    one run(list), (b) parallel="instances": instances dealt over the ranks, one all-reduce at the end, or
    parallel="trajectories": every rank walks the same loop and each circuit's shots are split over the ranks
    (dist.ShardedSampler: one Pauli trajectory per shot whatever L; counts identical to one GPU).  Seeds are `seed + 1000003 * instance + 4099 * step + evaluation index`,
    so the result does not depend on the number of ranks.

    Returns {"forward", "echo", "g_history": float64 [n_inst, T]; "mean_forward", "mean_echo", "mean_g": [T];
             "circuits": circuits evaluated over all ranks}."""
    from . import dist as D
    from .backend import compute_z_expectation
    if parallel not in ("instances", "trajectories"):
        raise ValueError("parallel must be 'instances' or 'trajectories'")
    hs = np.atleast_2d(np.asarray(hs, dtype=np.float64))
    phis = np.atleast_2d(np.asarray(phis, dtype=np.float64))
    n_inst = hs.shape[0]
    # "trajectories" always samples through the trajectory engine (also on one rank), so the result does not depend on the
    # number of ranks; "instances" goes through run(), whose automatic method takes the exact density matrix for small L
    split_shots = parallel == "trajectories"
    sampler = D.ShardedSampler(sim, rank, world, group) if split_shots else None
    mine = list(range(n_inst)) if (split_shots or world == 1) else D.deal_units(n_inst, rank, world)
    out = np.zeros((3, n_inst, T), dtype=np.float64)
    n_circ = 0

    def expvals(circs, seeds):
        if sampler is not None:
            return [compute_z_expectation(sampler.run_counts(c, shots, s), 1)[0] for c, s in zip(circs, seeds)]
        res = sim.run(circs, shots=shots, seed_simulator=[int(s) for s in seeds]).result()
        return [compute_z_expectation(res.get_counts(j), 1)[0] for j in range(len(circs))]

    for i in mine:
        g_hist = []
        g = float(g_initial)
        for t in range(T):
            gv = g_hist + [g]
            g_hist.append(g)
            base = int(seed_simulator) + 1000003 * i + 4099 * t
            circs = [autocorr_circuit(L, g, hs[i], phis[i], t + 1, qubit, echo, "x", initial_state, g_values=gv)
                     for echo in (False, True)]
            fwd, ech = expvals(circs, [base, base + 1])
            n_circ += 2
            out[0, i, t], out[1, i, t], out[2, i, t] = fwd, ech, g
            if t == T - 1:
                break
            if use_optimization:
                evals = [0]

                def echo_of(cands, _i=i, _t=t, _prev=g_hist[:-1], _base=base):
                    cs = [autocorr_circuit(L, c, hs[_i], phis[_i], _t + 1, qubit, True, "x", initial_state, g_values=_prev + [c])
                          for c in cands]
                    seeds = [_base + 2 + evals[0] + k for k in range(len(cs))]
                    evals[0] += len(cs)
                    return expvals(cs, seeds)

                g = optimize_g(echo_of, target_echo, g_min, g_max, optimizer, grid_points)
                n_circ += evals[0]
            else:
                g = feedback_g(ech, target_echo, g, t, feedback_gain, g_min, g_max, bool(exponential_feedback), decay_compensation)
    if not split_shots and world > 1:
        tot = D.all_reduce_sum(np.concatenate([out.reshape(-1), [float(n_circ)]]), group, sim.ctx.device)
        out, n_circ = tot[:-1].reshape(out.shape), int(round(tot[-1]))
    return {"forward": out[0], "echo": out[1], "g_history": out[2], "mean_forward": out[0].mean(axis=0),
            "mean_echo": out[1].mean(axis=0), "mean_g": out[2].mean(axis=0), "circuits": n_circ}


# ----------------------------------------------------------------------------------- energy sweeps (energy.py)
def energy_circuit(L, g, hs, phis, t, echo=False, initial_state="vacuum", transpile=True):
    """The L-qubit circuit of the energy scripts (energy.py:111-135): no ancilla, t periods of U_F on qubits 0..L-1, then
    -- echo -- t inverse periods.  ("neel" flips the qubits of even index >= 2 below L; the reference's own loop runs to
    index L, which its L-qubit circuit does not have, so only "vacuum" ever ran there.)"""
    circ = QuantumCircuit(L)
    if initial_state == "neel":
        for i in range(2, L, 2):
            circ.x(i)
    elif initial_state != "vacuum":
        raise ValueError(f"unknown initial state {initial_state!r}")
    sub = QuantumCircuit(L)
    for i in range(L):
        sub.rx(PI * g, i)
    for i in range(0, L - 1, 2):
        sub.rzz(float(phis[i]), i, i + 1)
    for i in range(1, L - 1, 2):
        sub.rzz(float(phis[i]), i, i + 1)
    for i in range(L):
        sub.rz(float(hs[i]), i)
    for _ in range(t):
        circ.append(sub, range(L))
    if echo:
        inv = sub.inverse()
        for _ in range(t):
            circ.append(inv, range(L))
    return lower_level0(circ) if transpile else circ


def run_energy_sweep(sim, L, g, hs, phis, t_values, echo=False, initial_state="vacuum", precision=None, seed_simulator=None,
                     rank=0, world=1, group=None):
    """<H>(t) / L of the kicked-Ising Hamiltonian along the Floquet evolution: the loops of the energy scripts
    (`get_single_out` / `get_instances`, energy.py:173-195; per-site energy as written to the CSV, energy.py:218-231).

    Every (instance, t) point is one estimator pub: `BackendEstimatorV2(sim).run([(energy_circuit, dtc_hamiltonian)])`
    (energy.py:166-171; Z / ZZ terms share one measurement circuit, the X terms another; 4096 shots per circuit at the
    default precision).  Points are dealt over the ranks, one all-reduce at the end.  The noise model is the simulator's; the
    reference transpiles at the preset manager's default level (energy.py:152-158), which merges single-qubit gates and so moves
    the noisy-gate sites -- no fixture pins that placement; circuits here are lowered at level 0 like the
    autocorrelation scripts'.
    Returns {"energy_per_site": float64 [n_inst, len(t_values)], "mean": [len(t_values)], "stds": same shape as
    energy_per_site (standard error of <H>/L from the shot statistics), "points", "circuits"}."""
    from . import dist as D
    from .estimator import BackendEstimatorV2, dtc_hamiltonian
    hs = np.atleast_2d(np.asarray(hs, dtype=np.float64))
    phis = np.atleast_2d(np.asarray(phis, dtype=np.float64))
    n_inst = hs.shape[0]
    pts = [(i, k) for i in range(n_inst) for k in range(len(t_values))]
    out = np.zeros((2, n_inst, len(t_values)), dtype=np.float64)
    n_circ = 0
    for p in D.deal_units(len(pts), rank, world):
        i, k = pts[p]
        opts = {} if seed_simulator is None else {"seed_simulator": int(seed_simulator) + 7919 * p}
        est = BackendEstimatorV2(sim, options=opts)
        circ = energy_circuit(L, g, hs[i], phis[i], int(t_values[k]), echo, initial_state)
        res = est.run([(circ, dtc_hamiltonian(L, g, phis[i], hs[i]))], precision=precision).result()[0]
        out[0, i, k] = float(res.data.evs) / L
        out[1, i, k] = float(res.data.stds) / L
        n_circ += res.metadata["circuits"]
    if world > 1:
        tot = D.all_reduce_sum(np.concatenate([out.reshape(-1), [float(n_circ)]]), group, sim.ctx.device)
        out, n_circ = tot[:-1].reshape(out.shape), int(round(tot[-1]))
    return {"energy_per_site": out[0], "mean": out[0].mean(axis=0), "stds": out[1], "points": len(pts), "circuits": n_circ}


# ----------------------------------------------------------------------------------- site-resolved <Z_i(t)> (dtc_qasm.py)
def expz_circuit(L, g, hs, phis, t, state="0"):
    """The circuit dtc_qasm.py sends through OpenQASM (dtc_qasm.py:70-91): L qubits, state "1" flips qubit L // 2, t periods
    of U_F, every qubit measured into its own classical bit."""
    circ = QuantumCircuit(L, L)
    if state == "1":
        circ.x(L // 2)
    elif state != "0":
        raise ValueError(f"unknown state {state!r}")
    for _ in range(t):
        for i in range(L):
            circ.rx(PI * g, i)
        for i in range(0, L - 1, 2):
            circ.rzz(float(phis[i]), i, i + 1)
        for i in range(1, L - 1, 2):
            circ.rzz(float(phis[i]), i, i + 1)
        for i in range(L):
            circ.rz(float(hs[i]), i)
    for i in range(L):
        circ.measure(i, i)
    return circ


def run_expz_sweep(sim, L, g, hs, phis, T, state="0", shots=1024, seed_simulator=1234, via_qasm=True, exact=False,
                   rank=0, world=1, group=None, chunk=32):
    """<Z_site(t)> for t = 1..T-1 and every site: `get_single_out` / `get_instances` of dtc_qasm.py (:123-160).

    Each point is one L-qubit circuit measured on all qubits; <Z_i> = (N0 - N1) / shots of classical bit i
    (`compute_z_expectation`, dtc_qasm.py:105-121).  via_qasm=True hands the simulator the OpenQASM-2 text the script writes
    (`expz_circuit(...).qasm()`, parsed by `ir.from_qasm2`), exactly the reference's hand-over format; exact=True returns the
    simulated probabilities' <Z_i> instead of the shot estimate (`Result.expectation_z`).  Points (instance, t) are dealt over
    the ranks, one all-reduce at the end; seeds are seed + global point index.
    Returns {"expz": float64 [n_inst, L, T-1] (the layout `savecsv` writes: row = (instance, site), column = t),
             "mean": [L, T-1], "points"}."""
    from . import dist as D
    from .backend import compute_z_expectation
    hs = np.atleast_2d(np.asarray(hs, dtype=np.float64))
    phis = np.atleast_2d(np.asarray(phis, dtype=np.float64))
    n_inst = hs.shape[0]
    pts = [(i, t) for i in range(n_inst) for t in range(1, T)]
    mine = D.deal_units(len(pts), rank, world)
    out = np.zeros((n_inst, L, max(T - 1, 0)), dtype=np.float64)
    for a in range(0, len(mine), chunk):
        idx = mine[a:a + chunk]
        circs = []
        for k in idx:
            i, t = pts[k]
            c = expz_circuit(L, g, hs[i], phis[i], t, state)
            circs.append(c.qasm() if via_qasm else c)
        if not circs:
            continue
        res = sim.run(circs, shots=shots, seed_simulator=[int(seed_simulator) + int(k) for k in idx]).result()
        for j, k in enumerate(idx):
            i, t = pts[k]
            out[i, :, t - 1] = res.expectation_z(j)[:L] if exact else compute_z_expectation(res.get_counts(j), L)
    if world > 1:
        out = D.all_reduce_sum(out.reshape(-1), group, sim.ctx.device).reshape(out.shape)
    return {"expz": out, "mean": out.mean(axis=0), "points": len(pts)}


# ----------------------------------------------------------------------------------- echo vs number of shots (shots.py)
def run_shots_sweep(sim, L, g, hs, phis, t_values, shot_numbers=(100, 1000, 10000, 100000, 1000000), echo=True,
                    polarization="x", qubit=None, initial_state="vacuum", seed_simulator=1234, rank=0, world=1, group=None):
    """Instance-averaged (echo) autocorrelation for several shot counts: the outer loop of shots.py (:49,:247-262; one
    independent sweep per entry of `shot_numbers`, each point one circuit with that many Pauli trajectories -- large counts
    are batched through the state buffer by run()).  Sweep k uses seeds seed + k * 10007 + point index.
    Returns {"shots": list, "mean": float64 [len(shot_numbers), len(t_values)] (the `av_autocorr_echo` column of each CSV),
             "autocorr": [len(shot_numbers), n_inst, len(t_values)], "periods"}."""
    means, full, periods = [], [], 0
    for k, n in enumerate(shot_numbers):
        res = run_sweep(sim, L, [g], hs, phis, t_values, echoes=(bool(echo),), polarizations=(polarization,), qubit=qubit,
                        initial_state=initial_state, shots=int(n), seed_simulator=int(seed_simulator) + 10007 * k,
                        rank=rank, world=world, group=group)
        full.append(res["autocorr"][0, 0, 0])
        means.append(res["mean"][0, 0, 0])
        periods += res["periods"]
    return {"shots": [int(n) for n in shot_numbers], "mean": np.asarray(means), "autocorr": np.asarray(full),
            "periods": periods}
