"""Travelling half of the plugin-level pins: the oracles (splice / conditioning gather / loss combination / diffusion loss) against
tests/golden/plugins.npz, minted from the LIVE reference by `python -m oracle.gen_golden_plugins` (build container).  Runs anywhere —
no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

from oracle import plugin_scenarios as PS

GOLD = os.path.join(os.path.dirname(__file__), "golden", "plugins.npz")


@pytest.fixture(scope="module")
def gold():
    assert os.path.isfile(GOLD), "tests/golden/plugins.npz missing: run `python -m oracle.gen_golden_plugins` in the build container"
    return np.load(GOLD)


@pytest.mark.parametrize("i", range(len(PS.SPLICE_CASES)))
def test_splice_oracle_matches_golden(gold, i):
    assert torch.equal(PS.oracle_splice(*PS.SPLICE_CASES[i]), torch.from_numpy(gold[f"splice_{i}"]))


@pytest.mark.parametrize("i", range(len(PS.CAUSAL_CASES)))
def test_gather_and_losses_match_golden(gold, i):
    r = PS.oracle_causal(*PS.CAUSAL_CASES[i])
    assert torch.equal(r["enc"], torch.from_numpy(gold[f"causal_{i}_enc"]))
    assert abs(float(r["lm_loss"]) - float(gold[f"causal_{i}_lm_loss"])) < 1e-5          # fp32 GEMM order may differ across hosts
    assert abs(float(r["loss"]) - float(gold[f"causal_{i}_loss"])) < 1e-4
    if r["u_enc"] is not None:
        assert torch.equal(r["u_enc"], torch.from_numpy(gold[f"causal_{i}_u_enc"]))
        assert r["null_ids"] == gold[f"causal_{i}_null_ids"].tolist()


@pytest.mark.parametrize("i", range(len(PS.SDHEAD_CASES)))
def test_diffusion_loss_oracle_matches_golden(gold, i):
    got = float(PS.oracle_sdhead(*PS.SDHEAD_CASES[i]))
    want = float(gold[f"sdhead_{i}"])
    assert abs(got - want) <= 1e-4 * abs(want), (got, want)      # a wrong branch is off by percents; 1e-4 absorbs host SIMD differences
