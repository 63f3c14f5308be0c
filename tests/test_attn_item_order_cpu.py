"""Item order of the persistent attention kernels (`dllm_attn_item_order`, host mirror of the device-side `decode_item`):
every (tile, head x batch) pair is handed out exactly once; a scheduling window spans at most win_heads consecutive heads (K / V of the
heads in flight stay L2-resident: a plain heaviest-first order over all heads made every K/V tile an HBM read, profiles/r02r_attn_fwd_persist.md);
inside a window the heavier tiles come first (causal triangle balance)."""
import ctypes

import pytest

from dreamllm_b200 import _lib


def _order(ntiles, n_hb, grid, descending):
    L = _lib.lib()
    t, h, wh = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    out = []
    for w in range(ntiles * n_hb):
        L.dllm_attn_item_order(w, ntiles, n_hb, grid, int(descending), ctypes.byref(t), ctypes.byref(h), ctypes.byref(wh))
        out.append((t.value, h.value))
    return out, wh.value


@pytest.mark.parametrize("ntiles,n_hb,grid", [(16, 256, 296), (16, 256, 148), (32, 20, 148), (1, 40, 40), (3, 160, 148), (5, 37, 296), (8, 1, 8),
                                              (2, 1000, 296), (64, 7, 148)])
@pytest.mark.parametrize("descending", [True, False])
def test_item_order_is_a_windowed_heaviest_first_permutation(ntiles, n_hb, grid, descending):
    items, win_heads = _order(ntiles, n_hb, grid, descending)
    assert sorted(items) == [(t, h) for t in range(ntiles) for h in range(n_hb)]          # bijection
    assert 1 <= win_heads <= n_hb and win_heads == min(n_hb, max(1, -(-2 * grid // ntiles)))
    wsz = win_heads * ntiles
    for w0 in range(0, len(items), wsz):
        win = items[w0:w0 + wsz]
        heads = sorted({h for _, h in win})
        assert heads == list(range(heads[0], heads[0] + len(heads))) and len(heads) <= win_heads      # consecutive heads only
        tiles = [t for t, _ in win]
        assert tiles == sorted(tiles, reverse=descending)                                  # heaviest tile first inside the window
        assert len(win) == len(heads) * ntiles                                             # every tile of those heads
