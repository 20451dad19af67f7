"""umma_gemm raw-A kernel: 3xTF32 vs 3xBF16 on the encoder / head shapes; CUDA events, L2 flushed."""
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
for (M, N, K) in ((43900, 1024, 256), (43900, 256, 1024), (43900, 256, 256), (43900, 288, 256), (33400, 512, 256), (33400, 256, 256)):
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') * 0.05; b = torch.randn(N, device='cuda')
    wh, wl = torch.empty_like(w), torch.empty_like(w); y = torch.empty(M, N, device='cuda')
    w16h = torch.empty((N, K), dtype=torch.bfloat16, device='cuda'); w16l = torch.empty_like(w16h)
    nat.check(lib.pn_split_tf32(w.data_ptr(), wh.data_ptr(), wl.data_ptr(), w.numel(), st), "split")
    nat.check(lib.pn_split_bf16(w.data_ptr(), w16h.data_ptr(), w16l.data_ptr(), w.numel(), st), "split16")
    r3 = lambda: nat.check(lib.pn_linear_tc_rawa(x.data_ptr(), wh.data_ptr(), wl.data_ptr(), b.data_ptr(), y.data_ptr(), N, M, N, K, st), "rawa")
    r16 = lambda: nat.check(lib.pn_linear_tc_bf16x3(x.data_ptr(), w16h.data_ptr(), w16l.data_ptr(), b.data_ptr(), y.data_ptr(), N, M, N, K, st), "b16")
    ref = torch.addmm(b.double(), x.double(), w.double().t())
    a = t(r3); e3 = float((y.double() - ref).abs().max() / ref.abs().max())
    c = t(r16); e16 = float((y.double() - ref).abs().max() / ref.abs().max())
    print(f"M={M} N={N} K={K}: 3xTF32 {a:.1f} us ({2*M*N*K/a/1e6:.0f} TF/s, err {e3:.1e})   3xBF16 {c:.1f} us ({2*M*N*K/c/1e6:.0f} TF/s, err {e16:.1e})  bytes/us {(M*K*4+M*N*4)/c/1e3:.0f} GB/s")
