"""Runs tests/native/mn_probe.cu on the GPU: MN-major operands (which of LBO / SBO strides the K groups) and the
tensor-memory lane layout of an M = 64 accumulator.   python tests/native/mn_probe.py"""
import ctypes as C, os, subprocess, sys, torch
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libmn_probe.so")
if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "mn_probe.cu")):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                           "-I", os.path.join(ROOT, "com_marl_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(HERE, "mn_probe.cu"), "-o", SO])
if not torch.cuda.is_available():
    print("built", SO); sys.exit(0)
lib = C.CDLL(SO)
lib.mn_probe.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p] * 2
g = torch.Generator().manual_seed(1)
for mode, M, N, K in ((0, 128, 64, 32), (0, 128, 128, 64), (1, 64, 32, 32), (1, 64, 64, 64)):
    if mode == 0:
        A = (torch.rand((128, K), generator=g) * 2 - 1).cuda(); B = (torch.rand((K, N), generator=g) * 2 - 1).cuda()
        ref = A.half().float() @ B.half().float()                     # [128][N]
    else:
        A = (torch.rand((K, 64), generator=g) * 2 - 1).cuda(); B = (torch.rand((N, K), generator=g) * 2 - 1).cuda()
        ref = A.half().float().t() @ B.half().float().t()             # [64][N]
    for variant in ((0, 1) if (len(sys.argv) > 1 and sys.argv[1] == 'both') else (0,)):
        D = torch.zeros((128, 128), device="cuda"); st = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = lib.mn_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, mode, variant, st.data_ptr(), None)
        torch.cuda.synchronize()
        D = D.cpu(); refc = ref.cpu()
        if mode == 0:
            err = (D[:, :N] - refc).abs().max().item()
            print(f"mode 0 (MN-major B) N={N} K={K} variant={variant} (0: LBO = K-group stride, SBO = N-group stride): rc={rc} status={int(st.item())} max err {err:.3e}")
        else:
            # find the lane each row landed in
            lanes = []
            for r in range(64):
                d = (D[:, :N] - refc[r][None, :]).abs().max(dim=1).values
                l = int(d.argmin()); lanes.append((l, float(d[l])))
            worst = max(e for _, e in lanes)
            nz = [i for i in range(128) if D[i].abs().max() > 0]
            print(f"mode 1 (MN-major A, M=64) N={N} K={K} variant={variant}: rc={rc} status={int(st.item())} worst row err {worst:.3e}; row->lane "
                  f"{[l for l, _ in lanes[:20]]}... rows 16,32,48 -> {lanes[16][0]}, {lanes[32][0]}, {lanes[48][0]}; nonzero lanes {nz[:4]}..{nz[-4:]} ({len(nz)})")
