"""bench.py contract on the CPU: the reference arm (oracle port on the host cores) prints ONE JSON line with the keys the
driver reads; workload bookkeeping (sweep points, period counts) matches BASELINE config C2."""
import argparse
import json

import bench


def test_c2_bookkeeping():
    pts = bench.sweep_points(30)
    assert len(pts) == 60 and pts[0] == (0, False) and pts[-1] == (29, True)
    assert bench.periods_of(pts) == 1305                      # 1305 periods per sweep
    cfg = bench.workload_config(argparse.Namespace(tmax=30, trajectories=1024, gpus=1))
    assert "workload" in cfg and "model" not in cfg and cfg["state_bytes_n21"] == 32 << 20


def test_reference_arm_json_line(monkeypatch, capsys):
    monkeypatch.setattr(bench, "CPU_SAMPLE", [(1, False), (1, True)])       # 3 periods instead of 510: seconds, not minutes
    monkeypatch.setattr(bench, "CPU_TRAJ", 1)
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=1, steps=1, warmup=1, tmax=30, trajectories=1024)
    bench.run_reference(args)
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "periods/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "periods/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2")


def test_reference_arm_only_on_rank_zero(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(argparse.Namespace(gpus=2, steps=1, warmup=1, tmax=30, trajectories=1024))
    assert capsys.readouterr().out.strip() == ""
