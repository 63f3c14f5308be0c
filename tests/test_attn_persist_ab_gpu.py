"""Persistent attention kernels == one-CTA-per-item kernels, bit for bit (scripts/attn_ab_check.py).

Both variants do the same arithmetic in the same order, so ANY difference is a synchronisation bug in the persistent kernels' cross-item
pipelining (operand prefetch, accumulator staging in the operand ring, item scheduler).  The check found exactly one in round 2: with
items short enough to fit the ring the producer ran two items ahead of the epilogues and a parity wait on the "stored" mbarrier
aliased (profiles/r02x_attn_bwd_race.md).  Shapes include the C2 attention shape at B = 4 (13.8 items per CTA) and the UNet / CLIP /
cross-attention / padded shapes; every shape runs three times."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_persistent_kernels_bit_equal_one_cta_per_item(tmp_path):
    ref = str(tmp_path / "attn_ref.pt")
    env = dict(os.environ, DLLM_ATTN_NONPERSIST="1")
    r = subprocess.run([sys.executable, "scripts/attn_ab_check.py", "save", ref], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    env = {k: v for k, v in os.environ.items() if k != "DLLM_ATTN_NONPERSIST"}
    r = subprocess.run([sys.executable, "scripts/attn_ab_check.py", "cmp", ref], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "bad 0" in r.stdout
