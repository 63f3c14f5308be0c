"""bench.py host logic: the algorithmic-work formula behind `step_roofline` equals SURVEY.md §8(d)'s figures."""
import importlib.util
import os


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_flops_per_step_matches_survey_8d():
    b = _bench()
    c2 = b.flops_per_step(8, 2048)                                   # C2 model: 675.9 TFLOP fwd+bwd per step
    assert abs(c2 / 1e12 - 675.9) < 0.05
    assert abs(c2 / 3 / (8 * 2048) / 1e9 - 13.751) < 0.001          # 13.751 GFLOP / token forward
    c1_layer = b.flops_per_step(1, 512, layers=1) - 3 * 512 * 2 * 4096 * 32008
    assert abs(c1_layer / 1e9 - 628.14) < 0.01                       # a7 at C1: 628.14 GFLOP fwd+bwd


def test_metric_and_workload_names_follow_baseline_json():
    import json
    b = _bench()
    base = json.load(open(os.path.join(os.path.dirname(__file__), "..", "BASELINE.json")))
    assert b.METRIC == base["metric"]
    import argparse
    cfg = b.workload_config(argparse.Namespace(layers=b.L, bs=8, seq=2048, gpus=1))
    assert "configs[1]" in cfg["workload"] and "model" not in cfg
