"""General single-qubit noise channels (thermal relaxation, amplitude / phase damping, Kraus sets, reset): the non-Pauli part
of a device-calibrated noise model (NoiseModel.from_backend, fast.py:77-78).  Exact on the density-matrix path;
the trajectory path refuses them.  CPU: constructors against closed forms, planner segments through the numpy interpreter and
the emulated kernels against the oracle's Kraus evolution."""
import math

import numpy as np
import pytest

import dtcsim
import emu
import program_interp as PI
import refcircuits as RC
from dtcsim import compile_circuit
from dtcsim import noise as N
from oracle import oracle as O


def _apply(S, rho):
    """rho (2 x 2) through a superoperator with index = row + 2 col."""
    return (S @ rho.reshape(-1, order="F")).reshape(2, 2, order="F")


def test_thermal_relaxation_closed_form():
    t1, t2, t, p1 = 100.0, 60.0, 7.0, 0.1
    e = N.thermal_relaxation_error(t1, t2, t, p1)
    assert isinstance(e, N.ChannelError)
    rho = np.array([[0.3, 0.2 - 0.1j], [0.2 + 0.1j, 0.7]])
    out = _apply(e.S, rho)
    pr = 1 - math.exp(-t / t1)
    assert out[0, 0] == pytest.approx((1 - pr) * 0.3 + pr * 0.9) and out[1, 1] == pytest.approx((1 - pr) * 0.7 + pr * 0.1)
    assert out[0, 1] == pytest.approx(math.exp(-t / t2) * (0.2 - 0.1j))
    # both of Aer's constructions (mixture for T2 <= T1, Choi matrix above) are this channel: the oracle's Kraus sets agree
    for tt2 in (60.0, 150.0):
        S = N.thermal_relaxation_error(t1, tt2, t, p1).S
        ks = O.thermal_relaxation_kraus(t1, tt2, t, p1)
        assert np.abs(sum(np.kron(np.conj(K), K) for K in ks) - S).max() < 1e-14
        assert np.abs(sum(K.conj().T @ K for K in ks) - np.eye(2)).max() < 1e-14
    # pure dephasing (T1 = inf) is a Pauli-Z mixture and stays on the Pauli path
    pz = N.thermal_relaxation_error(math.inf, 50.0, 5.0)
    assert isinstance(pz, N.PauliError) and pz.probs == pytest.approx((0, 0, (1 - math.exp(-0.1)) / 2))
    with pytest.raises(ValueError):
        N.thermal_relaxation_error(10.0, 25.0, 1.0)                     # T2 > 2 T1


def test_damping_constructors_and_composition():
    g = 0.3
    out = _apply(N.amplitude_damping_error(g).S, np.array([[0.0, 0.5], [0.5, 1.0]]))
    assert out[0, 0] == pytest.approx(g) and out[1, 1] == pytest.approx(1 - g) and out[0, 1] == pytest.approx(0.5 * math.sqrt(1 - g))
    pd = N.phase_damping_error(0.36)
    assert isinstance(pd, N.PauliError) and pd.pz == pytest.approx(0.1)
    # Pauli o channel composes to a channel; channel o its own inverse-free partner keeps trace preservation
    c = N._compose(N.depolarizing_error(0.1, 1), N.amplitude_damping_error(g))
    assert isinstance(c, N.ChannelError) and np.abs(c.S[0] + c.S[3] - np.array([1, 0, 0, 1])).max() < 1e-15
    # a Kraus set that is a Pauli mixture is recognised as one (runs on every method)
    e = N.kraus_error([math.sqrt(0.9) * np.eye(2), math.sqrt(0.1) * np.array([[0, 1], [1, 0]])])
    assert isinstance(e, N.PauliError) and e.probs == pytest.approx((0.1, 0, 0))
    with pytest.raises(ValueError):
        N.ChannelError(np.eye(4) * 0.5)                                  # not trace preserving


def test_noise_model_dict_with_reset_and_kraus():
    """to_dict() form of a device model entry: mixture of {id, z, reset} circuits (Aer's thermal relaxation for T2 <= T1) and a
    Kraus instruction, in qiskit's serialised [[re, im], ...] form."""
    t1, t2, t = 100.0, 60.0, 7.0
    pr = 1 - math.exp(-t / t1)
    pz = (1 - pr) * (1 - math.exp(-t * (1 / t2 - 1 / t1))) / 2
    d = {"errors": [{"type": "qerror", "operations": ["u3"], "gate_qubits": [[2]],
                     "instructions": [[{"name": "id", "qubits": [0]}], [{"name": "z", "qubits": [0]}], [{"name": "reset", "qubits": [0]}]],
                     "probabilities": [1 - pz - pr, pz, pr]},
                    {"type": "qerror", "operations": ["u2"],
                     "instructions": [[{"name": "kraus", "qubits": [0],
                                        "params": [[[[1, 0], [0, 0]], [[0, 0], [math.sqrt(0.75), 0]]],
                                                   [[[0, 0], [0.5, 0]], [[0, 0], [0, 0]]]]}]],
                     "probabilities": [1.0]}]}
    nm = dtcsim.as_noise_model(d)
    assert nm.has_channel_noise()
    assert np.abs(nm.lookup("u3", 2) - N.thermal_relaxation_error(t1, t2, t).S).max() < 1e-15
    assert nm.lookup("u3", 0) is None
    assert np.abs(nm.lookup("u2", 5) - N.amplitude_damping_error(0.25).S).max() < 1e-15


def _thermal_model(names=("u1", "u2", "u3")):
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(N.thermal_relaxation_error(80.0, 100.0, 4.0, 0.05), list(names))
    return nm, O.PauliNoise.thermal_relaxation(80.0, 100.0, 4.0, 0.05, names=names)


@pytest.mark.parametrize("L,t,echo", [(4, 2, True), (5, 1, False), (6, 1, True)])
def test_channel_segments_interpreter_and_emulator_vs_oracle(disorder, L, t, echo):
    """Hadamard-test circuit with thermal relaxation (T2 > T1: the Choi-matrix branch) after every u2 / u3: planner segments
    executed by the numpy interpreter and by the emulated kernels == the oracle's Kraus evolution of the gate list."""
    hs, phis = disorder[20][0][0][:L], disorder[20][1][0][:L - 1]
    circ = RC.transpiled(RC.qc_body("vacuum", L, 0.97, hs, phis, t, L // 2, echo))
    nm, onoise = _thermal_model()
    prog = compile_circuit(circ, nm, want_dm=True)
    assert prog.has_channels and any(s[0] == "K" for s in prog.dm_segments)
    oc, na, _ = O.compact_ops(RC.ops_of(circ), circ.num_qubits)
    want = O.run_density_matrix(oc, na, onoise)
    assert abs(np.trace(want) - 1) < 1e-12
    assert np.abs(PI.dm_to_circuit_order(PI.run_dm(prog, prog.n), prog) - want).max() < 1e-12
    rho = emu.dm_program(prog)                                            # [col, row]
    assert np.abs(PI.dm_to_circuit_order(rho.T, prog) - want).max() < 1e-12


def test_trajectory_path_refuses_channels(disorder):
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, 1, 2, False))
    nm, _ = _thermal_model()
    with pytest.raises(ValueError, match="density-matrix"):
        compile_circuit(circ, nm)
    with pytest.raises(ValueError):
        compile_circuit(circ, nm, want_dm=True, optimize=True)
    # a Pauli-only model built from channel constructors still compiles for trajectories
    nm2 = dtcsim.NoiseModel()
    nm2.add_all_qubit_quantum_error(N.phase_damping_error(0.1), ["u3"])
    assert compile_circuit(circ, nm2).n_sites > 0


def test_pauli_twirl_is_explicit_and_matches_closed_form(disorder):
    """Opt-in approximation: the Pauli twirl of thermal relaxation keeps the decay rates (pX = pY = p_reset / 4,
    pZ = (1 + lambda_z - 2 e2) / 4 for p1 = 0) and makes the model runnable on the trajectory path; nothing twirls silently."""
    t1, t2, t = 100.0, 120.0, 6.0
    e = N.thermal_relaxation_error(t1, t2, t)
    tw = e.pauli_twirl()
    pr, e2 = 1 - math.exp(-t / t1), math.exp(-t / t2)
    assert tw.px == pytest.approx(pr / 4) and tw.py == pytest.approx(pr / 4) and tw.pz == pytest.approx((2 - pr - 2 * e2) / 4)
    # the twirl of a Pauli mixture is the mixture itself
    d = N.ChannelError(N.depolarizing_error(0.08, 1).superop()).pauli_twirl()
    assert d.probs == pytest.approx((0.02, 0.02, 0.02))
    nm = dtcsim.NoiseModel()
    nm.add_all_qubit_quantum_error(e, ["u3"])
    nm.add_all_qubit_readout_error([[0.98, 0.02], [0.05, 0.95]])
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, 1, 2, False))
    with pytest.raises(ValueError):
        compile_circuit(circ, nm)                                        # exact model: trajectories refused
    tw_nm = nm.pauli_twirled()
    assert not tw_nm.has_channel_noise() and tw_nm.has_readout_noise() and nm.has_channel_noise()
    prog = compile_circuit(circ, tw_nm)
    assert prog.n_sites == 4 and not prog.has_channels


def test_errors_on_two_qubit_gates_are_refused(disorder):
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    circ = RC.transpiled(RC.qc_body("vacuum", 4, 0.84, hs, phis, 1, 2, False))
    for err in (dtcsim.depolarizing_error(0.01, 1), N.amplitude_damping_error(0.01)):
        nm = dtcsim.NoiseModel()
        nm.add_all_qubit_quantum_error(err, ["cx"])
        with pytest.raises(ValueError, match="multi-qubit gate"):
            compile_circuit(circ, nm, want_dm=True)
