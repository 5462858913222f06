import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from s3gaussian_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ok = True
for K, N in [(64, 64), (128, 64), (64, 48), (64, 16), (8, 16), (128, 32)]:
    A = torch.randn(128, K, device=dev, generator=g); B = torch.randn(N, K, device=dev, generator=g)
    ref = (A.double() @ B.double().t())
    for three in (0, 1):
        D = torch.full((128, N), float("nan"), device=dev)
        rc = lib.s3g_umma_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, three, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        err = float((D.double() - ref).abs().max() / ref.abs().max())
        print(f"K={K} N={N} three_pass={three} rc={rc} rel err {err:.3e}")
        ok &= (err < (2e-6 if three else 3e-3))
print("UMMA SELFTEST", "OK" if ok else "FAILED")

if "--mn" in sys.argv:
    # MN-major operands (round-2 building block): D = A^T B with the reduction over the 128 rows
    ok = True
    for N in (16, 32, 64):
        A = torch.randn(128, 128, device=dev, generator=g); B = torch.randn(128, N, device=dev, generator=g)
        ref = A.double().t() @ B.double()
        for three in (0, 1):
            D = torch.full((128, N), float("nan"), device=dev)
            rc = lib.s3g_umma_selftest_mn(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, three, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            err = float((D.double() - ref).abs().max() / ref.abs().max())
            print(f"MN-major N={N} three_pass={three} rc={rc} rel err {err:.3e}")
            ok &= (err < (2e-6 if three else 3e-3))
    print("UMMA MN-MAJOR SELFTEST", "OK" if ok else "FAILED")
