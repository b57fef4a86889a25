"""Regenerate tests/golden/ from /root/reference (run in the authoring container only).

The GPU box has no /root/reference, so everything tests need at run time is committed here:
  * disorder inputs               hs_L{4,20}.csv, phis_L{4,20}.csv        (verbatim copies, data files)
  * statistical anchors (1024 shots, sigma ~ 0.02-0.03), trimmed to the columns tests read
  * gate_counts.json              exact op multisets of the transpiled circuits (gate_counts_*.csv)
  * known_answers.json            oracle values; the L=4 ones equal SURVEY.md 8c digit for digit
Usage:  python tests/golden/make_golden.py
"""
import glob
import json
import os
import re
import shutil
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle as O, dtc_circuits as C  # noqa: E402


def copy_inputs():
    for f in ["hs_L4.csv", "phis_L4.csv", "hs_L20.csv", "phis_L20.csv"]:
        shutil.copy(os.path.join(REF, f), os.path.join(HERE, f))


def trim(src, dst, cols):
    df = pd.read_csv(src)
    df[[c for c in cols if c in df.columns]].to_csv(os.path.join(HERE, dst), index=False)


def copy_stats():
    d4 = os.path.join(REF, "autocorr_data_L4")
    for gain in ("0.01", "0.05"):
        src = glob.glob(os.path.join(d4, f"autocorr_data_vacuum_realtime_adaptive_g0.84_L4_*gain{gain}.csv"))[0]
        trim(src, f"ref_L4_adaptive_gain{gain}.csv",
             ["time", "av_autocorr_standard", "av_autocorr_echo_standard", "g_history_inst1",
              "echo_adaptive_inst1", "forward_adaptive_inst1"])
    dp = os.path.join(REF, "autocorr_data_L20_polarization")
    for pol in ("x", "y", "xy", "yx"):
        src = glob.glob(os.path.join(dp, f"autocorr_data_vacuum_g0.97_L20_*_pol{pol}_with_envelopes.csv"))[0]
        trim(src, f"ref_L20_pol_{pol}.csv", ["time", "av_autocorr", "av_autocorr_echo"])
    dc = os.path.join(REF, "autocorr_data_L20_circular-polarization")
    for pol in ("x", "y", "circular_left", "circular_right"):
        src = glob.glob(os.path.join(dc, f"autocorr_data_vacuum_g0.97_L20_*_pol{pol}_with_envelopes.csv"))[0]
        trim(src, f"ref_L20_circ_{pol}.csv", ["time", "av_autocorr", "av_autocorr_echo"])
    dg = os.path.join(REF, "controlled-autocorr_data_L20")
    src = glob.glob(os.path.join(dg, "autocorr_data_vacuum_realtime_adaptive_optimization_iter5_*.csv"))[0]
    trim(src, "ref_L20_controlled_iter5.csv",
         ["time", "g_history_inst1", "echo_adaptive_inst1", "forward_adaptive_inst1",
          "echo_standard_g84_inst1", "forward_standard_g84_inst1",
          "echo_standard_g97_inst1", "forward_standard_g97_inst1"])


def gate_counts():
    out = {}
    for key, folder in (("L4", "autocorr_data_L4"), ("L20_pol", "autocorr_data_L20_polarization"),
                        ("L20_circ", "autocorr_data_L20_circular-polarization"),
                        ("L20_ctrl", "controlled-autocorr_data_L20")):
        d = {}
        for f in sorted(glob.glob(os.path.join(REF, folder, "gate_counts_t*_aer_simulator_*.csv"))):
            m = re.search(r"gate_counts_t(\d+)_(forward|echo)_", os.path.basename(f))
            df = pd.read_csv(f)
            d[f"t{m.group(1)}_{m.group(2)}"] = {g: int(c) for g, c in zip(df["gate"], df["count"])}
        out[key] = d
    with open(os.path.join(HERE, "gate_counts.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)


def known_answers():
    hs = pd.read_csv(os.path.join(HERE, "hs_L4.csv")).values[0]
    phis = pd.read_csv(os.path.join(HERE, "phis_L4.csv")).values[0]
    ka = {"L4": {}}
    for g, p, ts in ((0.84, 0.05, list(range(0, 21))), (0.84, 0.0, [1, 2, 3, 4, 5]), (0.97, 0.05, [1, 2, 3])):
        noise = O.PauliNoise.depolarizing(p) if p else None
        rows = {}
        for t in ts:
            v = []
            for echo in (False, True):
                ops, n, nc = C.autocorr_gates("vacuum", 4, g, hs, phis, t, 2, echo)
                low = C.lower_level0(ops, C.SNAKE_LAYOUT)
                oc, na, _ = O.compact_ops(low, 31)
                rho = O.run_density_matrix(oc, na, noise)
                pr = O.outcome_probabilities(np.real(np.diag(rho)).copy(), na, O.measured_map(oc), 1)
                v.append(float(pr[0] - pr[1]))
            rows[str(t)] = v
        ka["L4"][f"g{g}_p{p}"] = rows
    # values quoted in SURVEY.md 8c (survey probe), kept verbatim as an independent pin
    ka["survey_8c"] = {
        "L4_g0.84_p0.05": {"0": [0.735091890625, 0.735091890625], "1": [-0.611957637491, 0.663420431289],
                           "2": [0.459808578785, 0.589576316167], "3": [-0.464011353702, 0.519480336946],
                           "4": [0.241187894719, 0.448920600020], "5": [-0.395014825237, 0.383964251348],
                           "10": [0.258314059073, 0.194431672765], "20": [0.177413055068, 0.055576693603]},
        "L4_g0.84_p0_forward": {"1": -0.876306680044, "2": 0.687503493101, "3": -0.741346609718,
                                "4": 0.366328197311, "5": -0.711802500799},
        "L4_g0.97_p0.05": {"1": [-0.695238050455, 0.663420431289], "2": [0.655287339066, 0.598388358974],
                           "3": [-0.623514561517, 0.539522661507]},
        "L20_g0.97_p0.05_Zq_forward": {"1": -0.945783866373, "2": 0.902387667645, "3": -0.853048706891,
                                       "4": 0.814049650340, "5": -0.768991818288},
        "L20_g0.97_p0.05_Zq_echo": {"1": 0.9025, "2": 0.814492484877, "3": 0.735066508176},
        "L20_g0.97_p0_Zq_forward": {"1": -0.995561964603, "2": 0.999894367391, "3": -0.994949236719},
    }
    with open(os.path.join(HERE, "known_answers.json"), "w") as fh:
        json.dump(ka, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    copy_inputs()
    copy_stats()
    gate_counts()
    known_answers()
    print("golden fixtures written to", HERE)
