"""Micro-benchmark (not a test): the two sort / scan implementations on rasteriser-shaped inputs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from wast3d_b200 import _lib

lib = _lib.load()


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    cases = []
    n = 22_600_000
    runs = ((torch.arange(n, device="cuda") % 11 + torch.randint(0, 4000, (n,), device="cuda", generator=g)) % 4346).to(torch.int32)
    cases.append(("tile ids (13 bits, runs)", runs, 0, 13))
    cases.append(("tile ids (13 bits, random)", torch.randint(0, 4346, (n,), device="cuda", generator=g, dtype=torch.int32), 0, 13))
    depth = (torch.rand(3_000_000, device="cuda", generator=g) * 20 + 0.2).view(torch.int32)
    cases.append(("depth keys (32 bits, 3M)", depth, 0, 32))
    for name, keys, b0, b1 in cases:
        vals = torch.arange(keys.numel(), device="cuda", dtype=torch.int32)
        ko, vo = torch.empty_like(keys), torch.empty_like(keys)
        for mode in (0, 1):
            t = timeit(lambda: _lib.check(lib.wast3d_test_sort_pairs(keys.numel(), keys.data_ptr(), vals.data_ptr(), ko.data_ptr(),
                                                                      vo.data_ptr(), b0, b1, mode, _lib.stream_ptr())))
            passes = (b1 - b0 + 7) // 8
            gb = keys.numel() * 16 * passes / 1e9
            print(f"{name:32s} mode {mode}: {t:.3f} ms  ({gb / t * 1e3:.0f} GB/s of pair traffic, {passes} passes)")
        ts = timeit(lambda: torch.sort(keys.to(torch.int64) & ((1 << (b1 - b0)) - 1), stable=True))
        print(f"{name:32s} torch.sort(stable) on int64: {ts:.3f} ms")
    x = torch.randint(0, 30, (3_000_000,), device="cuda", dtype=torch.int32)
    out = torch.empty_like(x); tot = torch.zeros(1, device="cuda", dtype=torch.int32)
    for mode in (0, 1):
        t = timeit(lambda: lib.wast3d_test_scan(x.numel(), x.data_ptr(), None, out.data_ptr(), tot.data_ptr(), mode, _lib.stream_ptr()))
        print(f"scan 3M mode {mode}: {t:.4f} ms")


if __name__ == "__main__":
    main()
