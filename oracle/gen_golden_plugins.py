"""Mint tests/golden/plugins.npz from the REFERENCE'S OWN CODE (build container only: needs /root/reference).

    python -m oracle.gen_golden_plugins

Scenarios and the verbatim exec of the reference methods live in oracle/plugin_scenarios.py.  Stored: the reference's spliced embeddings
(`DreamLLMModel.forward`), conditioning rows / losses (`DreamLLMForCausalMLM.forward`) and diffusion losses (`StableDiffusionHead.forward`,
all six option branches).  tests/test_golden_plugins.py checks the travelling oracles against this file wherever it runs."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import plugin_scenarios as PS  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "plugins.npz")


def main():
    assert PS.reference_available(), "needs /root/reference"
    d = {}
    for i, (n_images, with_dream) in enumerate(PS.SPLICE_CASES):
        out, _ = PS.live_splice(n_images, with_dream)
        d[f"splice_{i}"] = out.numpy()
    for i, (drop_prob, n_dm) in enumerate(PS.CAUSAL_CASES):
        r = PS.live_causal(drop_prob, n_dm)
        d[f"causal_{i}_enc"] = r["enc"].numpy()
        d[f"causal_{i}_lm_loss"] = np.float64(float(r["lm_loss"]))
        d[f"causal_{i}_loss"] = np.float64(float(r["loss"]))
        if r["u_enc"] is not None:
            d[f"causal_{i}_u_enc"] = r["u_enc"].numpy()
            d[f"causal_{i}_null_ids"] = np.asarray(r["null_ids"])
    for i, case in enumerate(PS.SDHEAD_CASES):
        d[f"sdhead_{i}"] = np.float64(float(PS.live_sdhead(*case)))
    np.savez_compressed(OUT, **d)
    print(f"wrote {OUT}: {len(d)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
