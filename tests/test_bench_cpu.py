"""bench.py host logic: the algorithmic-work formulas behind the rooflines equal SURVEY.md §8(d)'s figures, and the line names the
metric's own configuration."""
import argparse
import importlib.util
import json
import os


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_llm_flops_match_survey_8d():
    b = _bench()
    c2 = 3 * b.llm_fwd_flops(8 * 2048, 2048)                         # C2 model: 675.9 TFLOP fwd+bwd per step
    assert abs(c2 / 1e12 - 675.9) < 0.05
    assert abs(c2 / 3 / (8 * 2048) / 1e9 - 13.751) < 0.001          # 13.751 GFLOP / token forward
    c1_layer = 3 * b.llm_fwd_flops(512, 512, layers=1, with_head=False)
    assert abs(c1_layer / 1e9 - 628.14) < 0.01                       # a7 at C1: 628.14 GFLOP fwd+bwd


def test_c5_flops_are_llm_plus_vae_plus_unet():
    b = _bench()
    S = 1 + b.C5["txt"] + 1 + b.C5["Q"] + 1 + 1
    fl = b.c5_flops_per_gpu(b.C5["bs"], S)
    want = 2 * b.llm_fwd_flops(4 * S, S) + 4 * (1116.7e9 + 2 * 804.3e9)   # fwd + dgrad-only bwd; VAE fwd only
    assert fl == want and 15e12 < fl < 30e12


def test_metric_and_workload_names_follow_baseline_json():
    b = _bench()
    base = json.load(open(os.path.join(os.path.dirname(__file__), "..", "BASELINE.json")))
    assert b.METRIC == base["metric"]
    cfg = b.c5_config(argparse.Namespace(layers=b.L, bs=8, seq=2048, gpus=1), 1)
    assert "configs[4]" in cfg["workload"] and "model" not in cfg and cfg["seq_len"] == 100 and cfg["global_batch"] == 4
    assert b.C5["res"] == 512 and b.UNIT == "tokens+pixels/s"
