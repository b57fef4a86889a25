"""The oracle pinned against everything the reference offers for this path."""
import numpy as np
import pytest

from conftest import family_z, golden_csv
from oracle import c_oracle as CO
from oracle import dtc_circuits as C
from oracle import oracle as O
from oracle import philox


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def _l4_signal(disorder, g, p, t, echo, g_values=None):
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    ops, _, _ = C.autocorr_gates("vacuum", 4, g, hs, phis, t, 2, echo, g_values=g_values)
    low = C.lower_level0(ops, C.SNAKE_LAYOUT)
    oc, na, _ = O.compact_ops(low, 31)
    rho = O.run_density_matrix(oc, na, O.PauliNoise.depolarizing(p) if p else None)
    pr = O.outcome_probabilities(np.real(np.diag(rho)).copy(), na, O.measured_map(oc), 1)
    return pr[0] - pr[1]


def test_l4_known_answers_survey(disorder, known):
    """Exact 5-qubit density-matrix values quoted in SURVEY.md 8c, digit for digit."""
    s = known["survey_8c"]
    for t, (f, e) in s["L4_g0.84_p0.05"].items():
        assert abs(_l4_signal(disorder, 0.84, 0.05, int(t), False) - f) < 1e-11
        assert abs(_l4_signal(disorder, 0.84, 0.05, int(t), True) - e) < 1e-11
    for t, f in s["L4_g0.84_p0_forward"].items():
        assert abs(_l4_signal(disorder, 0.84, 0.0, int(t), False) - f) < 1e-11
        assert abs(_l4_signal(disorder, 0.84, 0.0, int(t), True) - 1.0) < 1e-11
    for t, (f, e) in s["L4_g0.97_p0.05"].items():
        assert abs(_l4_signal(disorder, 0.97, 0.05, int(t), False) - f) < 1e-11
        assert abs(_l4_signal(disorder, 0.97, 0.05, int(t), True) - e) < 1e-11


def test_l20_lightcone_known_answers(disorder, known):
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    s = known["survey_8c"]
    for t, v in s["L20_g0.97_p0.05_Zq_forward"].items():
        if int(t) <= 4:
            assert abs(O.lightcone_zq(20, 0.97, hs, phis, int(t), 10, 0.05) - v) < 1e-11
    for t, v in s["L20_g0.97_p0.05_Zq_echo"].items():
        if int(t) <= 2:
            assert abs(O.lightcone_zq(20, 0.97, hs, phis, int(t), 10, 0.05, echo=True) - v) < 1e-11
    for t, v in s["L20_g0.97_p0_Zq_forward"].items():
        assert abs(O.lightcone_zq(20, 0.97, hs, phis, int(t), 10, 0.0) - v) < 1e-11


def test_ancilla_factorisation_identity(disorder):
    """signal = 0.95^6 * <Z_q>: full 5-qubit DM vs light-cone chain value."""
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    for t in (1, 2):
        full = _l4_signal(disorder, 0.84, 0.05, t, False)
        zq = O.lightcone_zq(4, 0.84, hs, phis, t, 2, 0.05)
        assert abs(full - 0.95 ** 6 * zq) < 1e-12


def _chi2(model, data, shots=1024):
    model, data = np.asarray(model), np.asarray(data)
    sig = np.sqrt(np.maximum(1 - model ** 2, 1e-3) / shots)
    return float(np.mean(((data - model) / sig) ** 2))


def _zmax(model, data, shots=1024):
    """Largest |data - model| in units of the binomial standard error of a `shots`-shot estimate of `model`.
    Compared with conftest.family_z(N): the 3-sigma bar of BASELINE.json's north_star held family-wise over N points."""
    model, data = np.asarray(model), np.asarray(data)
    sig = np.sqrt(np.maximum(1 - model ** 2, 1e-3) / shots)
    return float(np.abs((data - model) / sig).max())


@pytest.mark.parametrize("gain", ["0.01", "0.05"])
def test_l4_committed_csv_statistics(disorder, gain):
    """Reference's own 1024-shot Aer output (autocorr_data_L4/*gain*.csv): chi^2/dof ~ 1 and every point within the
    family-wise 3-sigma bound (conftest.family_z over the 2T points of each column pair)."""
    df = golden_csv(f"ref_L4_adaptive_gain{gain}.csv")
    T = len(df)
    fwd = [_l4_signal(disorder, 0.84, 0.05, i + 1, False) for i in range(T)]      # row i <-> t = i+1 (ctrl-g.py:412-416)
    ech = [_l4_signal(disorder, 0.84, 0.05, i + 1, True) for i in range(T)]
    c = _chi2(fwd + ech, list(df["av_autocorr_standard"]) + list(df["av_autocorr_echo_standard"]))
    assert 0.5 < c < 1.6, c
    assert _zmax(fwd + ech, list(df["av_autocorr_standard"]) + list(df["av_autocorr_echo_standard"])) < family_z(2 * T)
    # adaptive columns: per-step g from the committed g_history (time-dependent-g circuits, ctrl-g.py:196-241)
    gh = list(df["g_history_inst1"])
    fa = [_l4_signal(disorder, 0.84, 0.05, i + 1, False, g_values=gh) for i in range(T)]
    ea = [_l4_signal(disorder, 0.84, 0.05, i + 1, True, g_values=gh) for i in range(T)]
    c2 = _chi2(fa + ea, list(df["forward_adaptive_inst1"]) + list(df["echo_adaptive_inst1"]))
    assert 0.5 < c2 < 1.6, c2
    assert _zmax(fa + ea, list(df["forward_adaptive_inst1"]) + list(df["echo_adaptive_inst1"])) < family_z(2 * T)


@pytest.mark.parametrize("pol", ["x", "y", "xy", "yx"])
def test_l20_polarization_csv_statistics(disorder, pol):
    """autocorr_data_L20_polarization/*pol{x,y,xy,yx}*.csv vs exact light-cone values (t small)."""
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    df = golden_csv(f"ref_L20_pol_{pol}.csv")
    tmax_f = 4 if pol in ("x", "y") else 2          # xy/yx have two kick layers per period
    tmax_e = 2 if pol in ("x", "y") else 1
    anc = 0.95 ** 6
    model, data = [], []
    for t in range(0, tmax_f + 1):
        model.append(anc * O.lightcone_zq(20, 0.97, hs, phis, t, 10, 0.05, polarization=pol))
        data.append(df["av_autocorr"][t])
    for t in range(0, tmax_e + 1):
        model.append(anc * O.lightcone_zq(20, 0.97, hs, phis, t, 10, 0.05, echo=True, polarization=pol))
        data.append(df["av_autocorr_echo"][t])
    z = (np.array(data) - np.array(model)) / np.sqrt(np.maximum(1 - np.array(model) ** 2, 1e-3) / 1024)
    assert np.abs(z).max() < family_z(len(z)), z
    assert np.mean(z ** 2) < 2.5, z


@pytest.mark.parametrize("pol", ["x", "y", "circular_left", "circular_right"])
def test_l20_circular_csv_statistics(disorder, pol):
    """autocorr_data_L20_circular-polarization/*pol{x,y,circular_left,circular_right}*.csv (circ-pol.py:110-173: per-step
    kick angles pi g cos(w k)/sqrt 2, +-pi g sin(w k)/sqrt 2 with w = --circular_frequency = 1.0, echo undoing the steps
    in reverse order) vs exact light-cone values with hs_L20/phis_L20 row 0."""
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    df = golden_csv(f"ref_L20_circ_{pol}.csv")
    two_layers = pol.startswith("circular")          # rx + ry per qubit per period: 2L noisy u3
    tmax_f, tmax_e = (3, 1) if two_layers else (4, 2)
    anc = 0.95 ** 6

    def period(step):
        return C.uf_gates(20, 0.97, phis, hs, pol, time_step=step, circular_frequency=1.0)

    model, data = [], []
    for t in range(0, tmax_f + 1):
        model.append(anc * O.lightcone_zq(20, 0.97, hs, phis, t, 10, 0.05, period_fn=period))
        data.append(df["av_autocorr"][t])
    for t in range(0, tmax_e + 1):
        model.append(anc * O.lightcone_zq(20, 0.97, hs, phis, t, 10, 0.05, echo=True, period_fn=period))
        data.append(df["av_autocorr_echo"][t])
    assert _zmax(model, data) < family_z(len(model)), (model, data)
    assert _chi2(model, data) < 2.5


def test_l20_controlled_g_csv_statistics(disorder):
    """controlled-autocorr_data_L20/*optimization_iter5*.csv (g-opt.py:530-545): row i is the circuit of i+1 periods whose
    step k uses g_history_inst1[k] (time-dependent g, echo undoing the steps in reverse order); the fixed-g baselines
    g = 0.84 and g = 0.97 (g-opt.py get_instances) sit in the *_standard_* columns.  hs_L20/phis_L20 row 0."""
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    df = golden_csv("ref_L20_controlled_iter5.csv")
    gh = [float(x) for x in df["g_history_inst1"]]
    anc = 0.95 ** 6
    model, data = [], []
    for i in range(4):                               # forward: i+1 periods, light cone of 2i+1 sites
        t = i + 1
        pf = lambda step: C.uf_gates(20, gh[step], phis, hs, "x")
        model.append(anc * O.lightcone_zq(20, None, hs, phis, t, 10, 0.05, period_fn=pf))
        data.append(df["forward_adaptive_inst1"][i])
        for g, col in ((0.84, "forward_standard_g84_inst1"), (0.97, "forward_standard_g97_inst1")):
            model.append(anc * O.lightcone_zq(20, g, hs, phis, t, 10, 0.05))
            data.append(df[col][i])
    for i in range(2):                               # echo: 2(i+1) periods
        t = i + 1
        pf = lambda step: C.uf_gates(20, gh[step], phis, hs, "x")
        model.append(anc * O.lightcone_zq(20, None, hs, phis, t, 10, 0.05, echo=True, period_fn=pf))
        data.append(df["echo_adaptive_inst1"][i])
        for g, col in ((0.84, "echo_standard_g84_inst1"), (0.97, "echo_standard_g97_inst1")):
            model.append(anc * O.lightcone_zq(20, g, hs, phis, t, 10, 0.05, echo=True))
            data.append(df[col][i])
    assert _zmax(model, data) < family_z(len(model)), (model, data)
    assert _chi2(model, data) < 2.0


def test_gate_counts_match_reference(disorder, gate_counts):
    """Op multisets of the lowered circuits == committed gate_counts_*.csv (exact structural pin)."""
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    for key, want in gate_counts["L4"].items():
        t, kind = key.split("_")
        ops, _, _ = C.autocorr_gates("vacuum", 4, 0.84, hs, phis, int(t[1:]), 2, kind == "echo")
        assert C.count_ops(C.lower_level0(ops, C.SNAKE_LAYOUT)) == want, key
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    for key in ("t1_forward", "t3_echo", "t19_echo"):
        t, kind = key.split("_")
        ops, _, _ = C.autocorr_gates("vacuum", 20, 0.97, hs, phis, int(t[1:]), 10, kind == "echo", polarization="yx")
        assert C.count_ops(C.lower_level0(ops, C.SNAKE_LAYOUT)) == gate_counts["L20_pol"][key], key
    for key in ("t1_forward", "t20_echo"):
        t, kind = key.split("_")
        ops, _, _ = C.autocorr_gates("vacuum", 20, 0.84, hs, phis, int(t[1:]), 10, kind == "echo")
        assert C.count_ops(C.lower_level0(ops, C.SNAKE_LAYOUT)) == gate_counts["L20_ctrl"][key], key
    ops, _, _ = C.autocorr_gates("vacuum", 20, 0.97, hs, phis, 29, 10, True, polarization="circular_left",
                                 circular_frequency=1.0)
    assert C.count_ops(C.lower_level0(ops, C.SNAKE_LAYOUT)) == gate_counts["L20_circ"]["t29_echo"]


def test_trajectory_mean_converges_to_density_matrix(disorder):
    """Pauli-trajectory average == exact channel (the equivalence Aer's two methods rely on)."""
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    ops, _, _ = C.autocorr_gates("vacuum", 4, 0.84, hs, phis, 2, 2, False)
    oc, na, _ = O.compact_ops(C.lower_level0(ops, C.SNAKE_LAYOUT), 31)
    noise = O.PauliNoise.depolarizing(0.05)
    psi = O.run_trajectories(oc, na, noise, 99, np.arange(6000))
    p = O.outcome_probabilities(np.abs(psi) ** 2, na, O.measured_map(oc), 1)
    ez = p[:, 0] - p[:, 1]
    exact = _l4_signal(disorder, 0.84, 0.05, 2, False)
    assert abs(ez.mean() - exact) < 4 * ez.std() / np.sqrt(len(ez))


def test_run_counts_contract(disorder):
    hs, phis = disorder[4][0][0], disorder[4][1][0]
    ops, nq, nc = C.autocorr_gates("vacuum", 4, 0.84, hs, phis, 1, 2, False)
    low = C.lower_level0(ops, C.SNAKE_LAYOUT)
    counts, info = O.run_counts(low, 31, 1, shots=1024, noise=O.PauliNoise.depolarizing(0.05), seed=1234)
    assert info["method"] == "density_matrix" and sum(counts.values()) == 1024
    ez = O.compute_z_expectation(counts, 1)[0]
    assert abs(ez - (-0.611957637491)) < 4 * np.sqrt((1 - 0.612 ** 2) / 1024)
    counts2, info2 = O.run_counts(low, 31, 1, shots=16, noise=O.PauliNoise.depolarizing(0.05), seed=1234)
    assert info2["method"] == "statevector" and sum(counts2.values()) == 16


def test_c_oracle_matches_numpy(disorder):
    hs, phis = disorder[20][0][0], disorder[20][1][0]
    L = 10
    ops, _, _ = C.autocorr_gates("neel", L, 0.97, hs[:L], phis[:L - 1], 3, L // 2, True, polarization="xy")
    oc, na, _ = O.compact_ops(C.lower_level0(ops, C.SNAKE_LAYOUT), 31)
    noise = O.PauliNoise.depolarizing(0.05)
    ref = O.run_trajectories(oc, na, noise, 5, [3, 4])
    for i, t in enumerate((3, 4)):
        got = CO.run_trajectory(oc, na, noise, 5, t)
        assert np.abs(ref[i] - got).max() < 1e-13
