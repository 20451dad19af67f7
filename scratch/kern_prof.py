"""One isolated launch set of each tcgen05 kernel for `ncu --set full -k regex:...` captures."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat, ops
lib = nat.load()
dev = 'cuda'
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
st = torch.cuda.current_stream().cuda_stream
if which in ('gemm', 'all'):
    M, N, K = 33400, 512, 256
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.05; b = torch.zeros(N, device=dev)
    wh, wl = torch.empty_like(w), torch.empty_like(w); y = torch.empty(M, N, device=dev)
    nat.check(lib.pn_split_tf32(w.data_ptr(), wh.data_ptr(), wl.data_ptr(), w.numel(), st), "split")
    for _ in range(3):
        nat.check(lib.pn_linear_tc_rawa(x.data_ptr(), wh.data_ptr(), wl.data_ptr(), b.data_ptr(), y.data_ptr(), N, M, N, K, st), "rawa")
    torch.cuda.synchronize()
if which in ('gemm16', 'gemm16_ffn2'):
    M, N, K = (43900, 1024, 256) if which == 'gemm16' else (43900, 256, 1024)
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.05; b = torch.zeros(N, device=dev)
    wh = torch.empty((N, K), dtype=torch.bfloat16, device=dev); wl = torch.empty_like(wh); y = torch.empty(M, N, device=dev)
    nat.check(lib.pn_split_bf16(w.data_ptr(), wh.data_ptr(), wl.data_ptr(), w.numel(), st), "split16")
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for _ in range(3):
        flush.add_(1.0)
        nat.check(lib.pn_linear_tc_bf16x3(x.data_ptr(), wh.data_ptr(), wl.data_ptr(), b.data_ptr(), y.data_ptr(), N, M, N, K, st), "b16")
    torch.cuda.synchronize()
if which in ('fa', 'all'):
    B, Nq, Nk = 2, 100, 16700
    q = torch.randn(B, Nq, 256, device=dev) * 0.3; k = torch.randn(B, Nk, 256, device=dev) * 0.3; v = torch.randn(B, Nk, 256, device=dev)
    E = torch.randn(B, Nq, 256, device=dev); Fl = torch.randn(B, 256, (Nk + 63) // 64 * 64, device=dev)
    bits, rowany = ops.attn_mask_bits(E, Fl, Nk)
    for _ in range(3):
        ops.mha_core_tc(q, k, v, bits, rowany)
    torch.cuda.synchronize()
if which in ('pair', 'all'):
    for N, Bm in ((100, 4096), (400, 1024)):
        g = torch.Generator(device="cpu").manual_seed(1234)
        s = F.normalize(torch.randn(Bm, N, 256, generator=g)).to(dev); o = F.normalize(torch.randn(Bm, N, 256, generator=g)).to(dev)
        plan = ops.PpnPlan(Bm, N, 100, dev)
        for _ in range(3):
            plan.run_embeds(s, o)
        torch.cuda.synchronize()
