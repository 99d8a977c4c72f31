import ctypes as C, os, subprocess, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "tests", "native", "libtc_probe.so")
lib = C.CDLL(SO)
lib.tc_ts_probe.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p] * 3
for N, K in ((64, 64), (128, 64)):
    g = torch.Generator().manual_seed(N + K)
    A = (torch.rand((128, K), generator=g) * 2 - 1).cuda(); B = (torch.rand((N, K), generator=g) * 2 - 1).cuda()
    ref = A.half().float() @ B.half().float().t()
    for mode in (0, 3, 5):
        res = []
        for reps in (1, 9, 33):
            Cout = torch.zeros((128, N), device="cuda"); cyc = torch.zeros(1, dtype=torch.int64, device="cuda"); st = torch.zeros(1, dtype=torch.int32, device="cuda")
            for _ in range(2):
                rc = lib.tc_ts_probe(A.data_ptr(), B.data_ptr(), Cout.data_ptr(), N, K, mode, reps, cyc.data_ptr(), st.data_ptr(), None)
            torch.cuda.synchronize()
            err = (Cout / reps - ref).abs().max().item()
            res.append((reps, int(cyc.item()), err, int(st.item()), rc))
        per = (res[2][1] - res[1][1]) / (24 * (K // 16))
        print(f"N={N} K={K} mode={mode}: {res}  -> {per:.1f} cycles/MMA")
