"""`ops._cached_weight`: a derived-layout cache entry must not be served to a DIFFERENT tensor that happens to reuse the old one's id,
data pointer and version (freed model -> new model at the same addresses)."""
import weakref

import torch

from dreamllm_b200 import ops


def test_cached_weight_is_tied_to_the_tensor_object():
    built = []

    def build(w):
        built.append(float(w.sum()))
        return w * 2

    a = torch.ones(4)
    key = ("unit", 12345)
    assert torch.equal(ops._cached_weight(key, a, build), a * 2) and len(built) == 1
    assert torch.equal(ops._cached_weight(key, a, build), a * 2) and len(built) == 1          # hit
    a.add_(1)                                                                                  # in-place write: version bump -> rebuild
    assert torch.equal(ops._cached_weight(key, a, build), a * 2) and len(built) == 2
    # another tensor with the same (ptr, version, device, dtype) tag: forge the entry a freed-and-reallocated parameter would find
    b = torch.full((4,), 7.0)
    ops._WCACHE[key] = ((b.data_ptr(), b._version, b.device, b.dtype), torch.zeros(4), weakref.ref(a))
    assert torch.equal(ops._cached_weight(key, b, build), b * 2) and len(built) == 3          # not the stale zeros
    # entries die with their tensor
    del b
    assert key not in ops._WCACHE
