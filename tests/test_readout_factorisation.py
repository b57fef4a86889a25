"""Read-out factorisation (plan.compile_circuit(optimize=True) + the per-trajectory read-out of
csrc/dtc_readout.cuh, executed on the CPU by tests/emul, and its numpy restatement tests/readout_ref.py):
device part emulated (tests/emul), reduced density matrix taken with numpy, compared with the oracle's
per-trajectory outcome probabilities of the *full* (L+1)-qubit circuit."""
import numpy as np
import pytest

import dtcsim
import emu
import program_interp as PI
import refcircuits as RC
import readout_ref as readout
from dtcsim import compile_circuit
from oracle import oracle as O


def _rdm(psi, n, bits):
    """[T, 2^n] -> [T, 2^k, 2^k] reduced density matrix on `bits` (index bit i = bits[i])."""
    T = psi.shape[0]
    k = len(bits)
    idx = np.arange(1 << n)
    sub = np.zeros(1 << n, dtype=np.int64)
    rest = np.zeros(1 << n, dtype=np.int64)
    pos = 0
    for b in range(n):
        if b in bits:
            sub |= ((idx >> b) & 1) << bits.index(b)
        else:
            rest |= ((idx >> b) & 1) << pos
            pos += 1
    M = np.zeros((T, 1 << k, 1 << (n - k)), dtype=np.complex128)
    M[:, sub, rest] = psi
    return M @ np.conj(np.transpose(M, (0, 2, 1)))


def _probs_factorised(prog, seed, trajs, engine=0):
    st, fx, fz, ph, npass = emu.run(prog, n_traj=len(trajs), traj_offset=int(trajs[0]), seed=seed, engine=engine)
    layers, k_of = PI.build_layers(prog)
    masks, rfx, rfz, rph = PI.frame_walk(prog, k_of, seed, trajs)
    assert np.array_equal(fx, rfx)
    rdm = _rdm(st, prog.n_main, list(prog.small["reg_bits"]))
    ref = readout.simulate_small(prog, rdm, masks, fx)
    dev = emu.readout_small(prog, rdm, masks, fx)          # the function the GPU kernel k_readout_small runs
    assert np.abs(dev - ref).max() < 1e-13
    return dev, npass


def _probs_oracle(circ, onoise, seed, trajs):
    oc, na, _ = O.compact_ops(RC.ops_of(circ), circ.num_qubits)
    psi = O.run_trajectories(oc, na, onoise, seed, trajs) if onoise else \
        np.repeat(O.run_statevector(oc, na)[None], len(trajs), 0)
    return O.outcome_probabilities(np.abs(psi) ** 2, na, O.measured_map(oc), 1)


@pytest.mark.parametrize("L,t,echo,pol,state,p", [
    (4, 0, False, "x", "vacuum", 0.05), (4, 3, True, "x", "neel", 0.05), (6, 2, False, "xy", "vacuum", 0.3),
    (11, 2, True, "x", "vacuum", 0.05), (12, 3, False, "y", "neel", 0.2), (13, 2, True, "yx", "vacuum", 0.05),
])
def test_factorised_readout_equals_full_circuit(disorder, L, t, echo, pol, state, p):
    hs, phis = disorder[20][0][4][:L], disorder[20][1][4][:L - 1]
    circ = RC.transpiled(RC.qc_body(state, L, 0.9, hs, phis, t, L // 2, echo, pol))
    prog = compile_circuit(circ, RC.noise_model(p), optimize=True)
    assert prog.small is not None and prog.n_main == prog.n - 1           # the ancilla left the register
    assert prog.n_exec_layers <= (2 if echo else 1) * t * (2 if len(pol) == 2 else 1) + 2
    trajs = np.arange(20, 52)
    got, npass = _probs_factorised(prog, 321, trajs)
    want = _probs_oracle(circ, O.PauliNoise.depolarizing(p), 321, trajs)
    assert np.abs(got - want).max() < 1e-12


def test_factorised_ideal_and_pass_count(disorder):
    L = 20
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, 29, 10, True))
    prog = compile_circuit(circ, RC.noise_model(), optimize=True)
    assert prog.n_main == 20 and prog.n_exec_layers == 59
    rows, n = emu.schedule(prog)
    assert n == 59                                     # one state sweep per period (+1): R|A, then R|B D R'|B, ...
    assert (rows[:, 6] == 0).all()                     # every ZZ bond sits in one of the two phase tables


def test_other_circuits_are_left_alone():
    c = dtcsim.QuantumCircuit(4, 4)
    c.h(0)
    for q in range(3):
        c.cx(q, q + 1)
    c.measure_all()
    prog = compile_circuit(c, None, optimize=True)
    assert prog.small is None and prog.n_main == prog.n == 4
    c2 = dtcsim.QuantumCircuit(3, 1)                  # measured qubit entangled in the bulk: nothing to eliminate
    c2.h(0); c2.rx(0.3, 1); c2.rx(0.4, 2); c2.cz(0, 1); c2.rx(0.2, 0); c2.cz(0, 2); c2.rx(0.1, 2); c2.cz(1, 2)
    c2.rx(0.5, 0); c2.cz(0, 1); c2.measure(0, 0)
    prog2 = compile_circuit(c2, None, optimize=True)
    oc, na, _ = O.compact_ops(RC.ops_of(c2), 3)
    if prog2.small is None:
        psi = PI.to_circuit_order(PI.run(prog2), prog2)[0]
        assert np.abs(psi - O.run_statevector(oc, na)).max() < 1e-13
    else:
        got, _ = _probs_factorised(prog2, 0, np.arange(1))
        want = O.outcome_probabilities(np.abs(O.run_statevector(oc, na)) ** 2, na, O.measured_map(oc), 1)
        assert np.abs(got[0] - want).max() < 1e-12
