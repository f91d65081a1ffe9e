"""GPU: the binning primitives (stable radix sort of pairs, exclusive scan) that replace the
reference's CUB calls (rasterizer_impl.cu:279,305-310), both implementations, against torch."""
import pytest
import torch

from wast3d_b200 import _lib

pytestmark = pytest.mark.gpu


def _sort(keys, vals, b0, b1, mode):
    lib = _lib.load()
    n = keys.numel()
    ko, vo = torch.empty_like(keys), torch.empty_like(keys)
    st = lib.wast3d_test_sort_pairs(n, keys.data_ptr() if n else None, vals.data_ptr() if vals is not None and n else None,
                                    ko.data_ptr(), vo.data_ptr(), b0, b1, mode, _lib.stream_ptr())
    _lib.check(st, "test_sort_pairs")
    return ko, vo


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,b0,b1,kind", [
    (1, 0, 32, "rand"), (31, 0, 8, "rand"), (4096, 0, 13, "rand"), (4097, 0, 13, "rand"), (8191, 3, 16, "rand"),
    (100_003, 0, 32, "rand"), (1_000_000, 0, 13, "few"), (1_000_001, 0, 32, "equal"), (3_000_000, 0, 32, "float"),
    (5_000_000, 0, 13, "runs"), (777_777, 0, 21, "rand")])
def test_sort_pairs_stable(built, n, b0, b1, kind, mode):
    g = torch.Generator(device="cuda").manual_seed(n)
    if kind == "rand":
        keys = torch.randint(0, 2**31, (n,), device="cuda", generator=g, dtype=torch.int64)
    elif kind == "few":
        keys = torch.randint(0, 7, (n,), device="cuda", generator=g, dtype=torch.int64)
    elif kind == "equal":
        keys = torch.full((n,), 0x5A5A5A5A, device="cuda", dtype=torch.int64)
    elif kind == "float":   # positive float bits, like the depth keys
        keys = (torch.rand(n, device="cuda", generator=g) * 20 + 0.2).view(torch.int32).to(torch.int64)
    else:                   # runs of consecutive ids, like a Gaussian's tiles
        keys = (torch.arange(n, device="cuda") % 11 + torch.randint(0, 4000, (n,), device="cuda", generator=g)) % 4346
    keys = keys.to(torch.int32).contiguous()
    vals = torch.randint(0, 2**31 - 1, (n,), device="cuda", generator=g, dtype=torch.int32) if n % 2 else None
    ko, vo = _sort(keys, vals, b0, b1, mode)
    mask = ((1 << (b1 - b0)) - 1)
    digit = (keys.to(torch.int64) >> b0) & mask
    order = torch.sort(digit, stable=True).indices
    assert torch.equal(ko, keys[order])
    expect_vals = (vals if vals is not None else torch.arange(n, device="cuda", dtype=torch.int32))[order]
    assert torch.equal(vo, expect_vals)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 100_000, 3_000_000])
def test_scan(built, n, mode):
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randint(0, 50, (n,), device="cuda", generator=g, dtype=torch.int32)
    perm = torch.randperm(n, device="cuda", generator=g).to(torch.int32) if n % 2 else None
    out = torch.empty_like(x)
    total = torch.zeros(1, device="cuda", dtype=torch.int32)
    lib = _lib.load()
    st = lib.wast3d_test_scan(n, x.data_ptr(), perm.data_ptr() if perm is not None else None, out.data_ptr(),
                              total.data_ptr(), mode, _lib.stream_ptr())
    _lib.check(st, "test_scan")
    src = x[perm.long()] if perm is not None else x
    inc = torch.cumsum(src.to(torch.int64), 0)
    assert torch.equal(out.to(torch.int64), inc - src)
    assert total.item() == inc[-1].item()
