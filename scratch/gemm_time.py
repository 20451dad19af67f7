"""umma_gemm raw-A kernel: TMA-store epilogue on/off, several shapes; CUDA events, L2 flushed."""
import sys, statistics, torch
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
def t(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.mean(ts) * 1e3
for (M, N, K) in ((33400, 512, 256), (43900, 1024, 256), (43900, 256, 1024), (43900, 256, 256), (8400, 512, 256)):
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') * 0.05; b = torch.randn(N, device='cuda')
    wh, wl = torch.empty_like(w), torch.empty_like(w); y = torch.empty(M, N, device='cuda')
    nat.check(lib.pn_split_tf32(w.data_ptr(), wh.data_ptr(), wl.data_ptr(), w.numel(), st), "split")
    run = lambda: nat.check(lib.pn_linear_tc_rawa(x.data_ptr(), wh.data_ptr(), wl.data_ptr(), b.data_ptr(), y.data_ptr(), N, M, N, K, st), "rawa")
    res = {}
    for opt in (1, 0):
        lib.pn_set_option(nat.PN_OPT_UMMA_TMA_STORE, opt)
        us = t(run)
        res[opt] = us
        if opt == 1:
            ref = torch.addmm(b.double(), x.double(), w.double().t())
            err = float((y.double() - ref).abs().max() / ref.abs().max())
    lib.pn_set_option(nat.PN_OPT_UMMA_TMA_STORE, 1)
    print(f"M={M} N={N} K={K}: tma-store {res[1]:.1f} us ({2*M*N*K/res[1]/1e6:.0f} TFLOP/s)  stg {res[0]:.1f} us ({2*M*N*K/res[0]/1e6:.0f} TFLOP/s)  err {err:.2e}")
