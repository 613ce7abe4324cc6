"""Host logic of the integer shadows of carriers (kernels.py): which tensors may take the integers back without the
divide-and-round pass, and which must not.  Pure bookkeeping -- runs on CPU tensors, no kernels involved."""
import gc

import torch

import ivit_b200.kernels as K


def _carrier(q, s):
    c = q.to(torch.float32) * s                       # stands for the kernel that creates the carrier
    K._shadow_register(c, q, s)
    return c


def test_views_of_a_carrier_map_to_the_same_elements_of_the_integers():
    q = torch.arange(2 * 5 * 12, dtype=torch.int16).reshape(2, 5, 12)
    s = torch.tensor([0.5])
    c = _carrier(q, s)
    assert K._shadow_lookup(c, s) is q
    v = c.reshape(2, 5, 3, 4).permute(2, 0, 3, 1)     # qkv-style split + permute
    got = K._shadow_lookup(v, s.reshape(-1))
    assert got is not None and torch.equal(got, q.reshape(2, 5, 3, 4).permute(2, 0, 3, 1))
    sl = c[:, 1:, 4:8].transpose(-2, -1)              # slicing + transpose
    assert torch.equal(K._shadow_lookup(sl, s), q[:, 1:, 4:8].transpose(-2, -1))
    assert torch.equal(K._shadow_lookup(c[1], s), q[1])


def test_copies_other_scales_and_modified_carriers_have_no_shadow():
    q = torch.randint(-100, 100, (4, 8), dtype=torch.int8)
    s = torch.tensor([0.25])
    c = _carrier(q, s)
    assert K._shadow_lookup(c * 1.0, s) is None                       # arithmetic on the carrier: a new tensor
    assert K._shadow_lookup(c.t().contiguous(), s) is None            # a copy
    assert K._shadow_lookup(torch.cat([c, c]), s) is None
    assert K._shadow_lookup(c, torch.tensor([0.25])) is None          # equal value, but not the scale it was made with
    assert K._shadow_lookup(c.double(), s) is None                    # dtype changed
    c.add_(1.0)                                                       # in-place write invalidates (version counter)
    assert K._shadow_lookup(c, s) is None
    c2 = _carrier(q, s)
    s.mul_(2.0)                                                       # the scale changed in place
    assert K._shadow_lookup(c2, s) is None


def test_shadow_dies_with_the_carrier():
    q = torch.zeros(3, 3, dtype=torch.int32)
    s = torch.tensor([1.0])
    c = _carrier(q, s)
    key = id(c)
    assert key in K._SHADOW
    v = c[0]                                           # a view keeps the base (and its shadow) alive
    del c
    gc.collect()
    assert K._shadow_lookup(v, s) is not None
    del v
    gc.collect()
    assert key not in K._SHADOW


def test_integers_at_a_storage_offset_are_not_registered():
    """a contiguous slice of a bigger integer buffer (q.storage_offset() != 0) must not become a shadow --
    views of the carrier are re-derived with the carrier's absolute storage offsets, which would address other rows."""
    big = torch.arange(6 * 12, dtype=torch.int16).reshape(6, 12)
    q = big[2:4]                                       # contiguous, storage offset 24
    assert q.is_contiguous() and q.storage_offset() == 24
    s = torch.tensor([0.5])
    c = _carrier(q, s)
    assert K._shadow_lookup(c, s) is None and K._shadow_lookup(c[1], s) is None
    q0 = q.clone()                                     # offset 0: registered, views address the right rows
    c0 = _carrier(q0, s)
    assert torch.equal(K._shadow_lookup(c0[1], s), q0[1])


def test_host_scalar_cache_is_keyed_on_the_tensor_object():
    """a temporary scale tensor freed and reallocated at the same address with version 0 must not hit."""
    from ivit_b200.quantization_utils.ops import _host_scalar
    cache = {}
    a = torch.tensor([0.25])
    assert float(_host_scalar(a, cache)) == 0.25
    assert _host_scalar(a, cache) is cache["val"]
    b = torch.tensor([0.5])                            # another object (possibly the same address after `del a`)
    assert float(_host_scalar(b, cache)) == 0.5
    b.mul_(2.0)
    assert float(_host_scalar(b, cache)) == 1.0        # in-place change bumps the version
