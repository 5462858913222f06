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

if "--probe" in sys.argv:
    # Descriptor probe (csrc/umma_probe.cu): which shared-memory word does the tensor core read as A(m, k) / B(n, k)
    # for a given (major-ness, layout type, LBO, SBO)?  The other operand is a K-major identity, the probed operand's
    # image holds its own word index (two runs: index % 1024 and index // 1024, both exact in tf32).
    import numpy as np
    fn = lib.s3g_umma_probe
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_uint] * 9 + [C.c_void_p]

    def canon(rows, K):          # canonical K-major no-swizzle index map [rows][K] -> word
        r = np.arange(rows)[:, None]; k = np.arange(K)[None, :]
        return (r // 8) * (K // 4) * 32 + (k // 4) * 32 + (r % 8) * 4 + (k % 4)

    def idesc(M, N, a_mn, b_mn):
        return (1 << 4) | (2 << 7) | (2 << 10) | ((N >> 3) << 17) | ((M >> 4) << 24) | (int(a_mn) << 15) | (int(b_mn) << 16)

    def run(a_img, b_img, N, idsc, ad, bd, ksteps=1):
        a = torch.from_numpy(a_img.astype(np.float32)).to(dev); b = torch.from_numpy(b_img.astype(np.float32)).to(dev)
        D = torch.full((128, N), float("nan"), device=dev)
        rc = fn(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), D.data_ptr(), N, ksteps, idsc, *ad, *bd,
                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return rc, D.cpu().numpy()

    WORDS = 16384
    widx = np.arange(WORDS)
    rows = [0, 1, 2, 3, 4, 7, 8, 9, 16, 32, 64, 127]
    # identity operand: 8 x 8 (K = 8), stored K-major canonical, for N = 8 / as A with M = 128 rows
    idB = np.zeros(8 * 8); idB[canon(8, 8)[np.arange(8), np.arange(8)]] = 1.0
    idA = np.zeros(128 * 8); idA[canon(128, 8)[np.arange(8), np.arange(8)]] = 1.0
    kmaj = (128, 256, 0, 256)            # lbo, sbo, layout, step of a K-major [.][8] canonical tile
    print("== control: A K-major [128][8] = word index, B identity")
    img = np.zeros(128 * 8); img[canon(128, 8).reshape(-1)] = (np.arange(128 * 8) % 1024)
    rc, D = run(img, idB, 8, idesc(128, 8, 0, 0), kmaj, kmaj)
    print(" rc", rc, "rows 0,1,9:", D[0].tolist(), D[1].tolist(), D[9].tolist())
    for name, amn, layouts in (("A MN-major", 1, (0, 2, 4, 6, 1)),):
        for layout in layouts:
            for lbo, sbo in ((4096, 128), (128, 4096), (1024, 128), (128, 1024), (256, 128), (128, 256), (512, 1024), (1024, 512)):
                res = []
                for part in (widx % 1024, widx // 1024):
                    rc, D = run(part, idB, 8, idesc(128, 8, amn, 0), (lbo, sbo, layout, 0), kmaj)
                    res.append(D)
                addr = res[0] + 1024 * res[1]
                ok = np.isfinite(addr).all()
                print(f"== {name} layout={layout} lbo={lbo} sbo={sbo} rc={rc} finite={ok} nonzero={int((addr != 0).sum())}")
                if ok and (addr != 0).any():
                    for m in rows:
                        print(f"   m={m:3d} k0..7 -> words", [int(v) for v in addr[m]])
    print("== B MN-major probes (A = K-major identity in rows 0..7, D[m][n] = B(n, k=m))")
    for layout in (0, 2, 4, 6):
        for lbo, sbo in ((4096, 128), (128, 4096), (1024, 128), (128, 1024), (256, 128), (128, 256)):
            res = []
            for part in (widx % 1024, widx // 1024):
                rc, D = run(idA, part, 32, idesc(128, 32, 0, 1), kmaj, (lbo, sbo, layout, 0))
                res.append(D)
            addr = res[0] + 1024 * res[1]
            ok = np.isfinite(addr).all()
            print(f"== B MN-major layout={layout} lbo={lbo} sbo={sbo} rc={rc} finite={ok} nonzero={int((addr != 0).sum())}")
            if ok and (addr != 0).any():
                for k in range(8):
                    print(f"   k={k} n0..31 -> words", [int(v) for v in addr[k]])
