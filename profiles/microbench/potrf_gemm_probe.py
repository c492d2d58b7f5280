import sys, torch
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops
def timed(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
g = torch.Generator().manual_seed(0)
def spd(T):
    A = torch.randn(T, T, generator=g, dtype=torch.float64)
    return (A @ A.t() / T + torch.eye(T, dtype=torch.float64)).cuda()
for T in (128, 256, 512, 1024, 2048):
    A = spd(T); w = A.clone()
    def f():
        w.copy_(A); ops.potrf_big(w)
    tc = timed(lambda: w.copy_(A))
    print("potrf T=%d: %.1f us (copy %.1f)" % (T, timed(f) - tc, tc))
for (M, N, K) in ((1920, 128, 128), (1024, 128, 128), (1920, 1920, 128), (16000, 128, 128), (16000, 128, 384), (8192, 8192, 128), (8192, 8192, 256), (8192, 8192, 512), (16000, 16000, 128), (16000, 16000, 256), (16000,16000,512)):
    a = torch.randn(M, K, device="cuda", dtype=torch.float64); b = torch.randn(N, K, device="cuda", dtype=torch.float64)
    c = torch.zeros(M, N, device="cuda", dtype=torch.float64)
    t = timed(lambda: ops.gemm_nt(a, b, alpha=-1.0, beta=1.0, C=c), n=20)
    print("gemm %dx%dx%d: %.1f us  %.2f TF/s" % (M, N, K, t, 2.0 * M * N * K / t / 1e6))
