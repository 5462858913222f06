"""Developer check (gpurun): the standalone radix sort vs torch.sort(stable=True)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from s3gaussian_b200 import _lib


def sort_pairs(keys, vals, b0, b1):
    lib = _lib.load()
    n = keys.numel()
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    tmp = torch.empty(lib.s3g_sort_temp_bytes(n), dtype=torch.uint8, device=keys.device)
    kin, vin = keys.clone(), vals.clone()
    _lib.check(lib.s3g_sort_pairs_u32(n, kin.data_ptr(), vin.data_ptr(), ko.data_ptr(), vo.data_ptr(),
                                      b0, b1, tmp.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "sort")
    return ko, vo


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    ok = True
    for n, b0, b1, hi in [(1, 0, 32, 2**31), (31, 0, 8, 200), (4096, 0, 32, 2**31), (4097, 0, 14, 9600),
                          (100_000, 0, 32, 2**31), (1_000_003, 0, 14, 9600), (8_000_000, 0, 14, 9600),
                          (2_000_000, 0, 32, 2**31), (3_000_000, 4, 20, 2**24)]:
        keys = torch.randint(0, hi, (n,), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
        vals = torch.arange(n, device=dev, dtype=torch.int32)
        ko, vo = sort_pairs(keys, vals, b0, b1)
        torch.cuda.synchronize()
        mask = ((1 << (b1 - b0)) - 1)
        dig = (keys.to(torch.int64) >> b0) & mask
        _, perm = torch.sort(dig, stable=True)
        good = bool((vo.to(torch.int64) == perm).all()) and bool((ko == keys[perm]).all())
        print(f"n={n} bits[{b0},{b1}) -> {'OK' if good else 'MISMATCH'}")
        ok &= good
    # timing of the two production shapes
    for n, b1, hi in [(2_000_000, 32, 2**31), (8_000_000, 14, 9600)]:
        keys = torch.randint(0, hi, (n,), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
        vals = torch.arange(n, device=dev, dtype=torch.int32)
        for _ in range(3):
            sort_pairs(keys, vals, 0, b1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record()
        for _ in range(10):
            sort_pairs(keys, vals, 0, b1)
        e.record(); torch.cuda.synchronize()
        print(f"sort n={n} bits={b1}: {s.elapsed_time(e) / 10:.3f} ms (incl. 2 clones + temp alloc)")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
