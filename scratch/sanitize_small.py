"""Small instances of every tcgen05 / TMA kernel, for `compute-sanitizer --tool memcheck|racecheck|synccheck`."""
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat, ops
from oracle.head import OConvTiny
lib = nat.load()
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
torch.manual_seed(0)
if which in ('pair', 'all'):
    for B, N in ((12, 100), (11, 200), (3, 400)):
        s = F.normalize(torch.randn(B, N, 256), dim=-1).cuda(); o = F.normalize(torch.randn(B, N, 256), dim=-1).cuda()
        imp, idx, sp, op = ops.PpnPlan(B, N, 100, 'cuda').run_embeds(s, o)
        torch.cuda.synchronize()
        assert float((imp - torch.matmul(s, o.transpose(1, 2))).abs().max()) < 1e-5
    print('pair_topk ok')
if which in ('conv', 'all'):
    m = OConvTiny(mid_channels=64).cuda()
    x = torch.tanh(torch.randn(2, 40, 40, device='cuda'))
    y = ops.conv_tiny(x, m); torch.cuda.synchronize()
    with torch.no_grad():
        assert float((y - m(x)).abs().max()) < 1e-3
    print('conv2_umma ok')
if which in ('gemm', 'all'):
    x = torch.randn(1500, 256, device='cuda'); w = torch.randn(384, 256, device='cuda') * 0.05; b = torch.randn(384, device='cuda')
    y = ops.linear_tc(x, w, b); torch.cuda.synchronize()
    assert float((y - torch.addmm(b, x, w.t())).abs().max()) < 1e-3
    print('umma_gemm ok')
if which in ('fa', 'all'):
    B, Nq, Nk = 1, 100, 1200
    q = torch.randn(B, Nq, 256, device='cuda') * 0.3; k = torch.randn(B, Nk, 256, device='cuda') * 0.3; v = torch.randn(B, Nk, 256, device='cuda')
    E = torch.randn(B, Nq, 256, device='cuda'); Fl = torch.randn(B, 256, (Nk + 63) // 64 * 64, device='cuda')
    bits, rowany = ops.attn_mask_bits(E, Fl, Nk)
    ops.mha_core_tc(q, k, v, bits, rowany); torch.cuda.synchronize()
    print('fa_umma ok')

if which in ('r02v', 'all'):
    # round-2 final additions: bf16 pair + top-k, 3xBF16 GEMM (split rings, 8 epilogue warps), PDL chain (small head forward)
    for B, N in ((12, 100), (5, 200), (3, 400)):
        s = F.normalize(torch.randn(B, N, 256), dim=-1).to(torch.bfloat16).cuda()
        o = F.normalize(torch.randn(B, N, 256), dim=-1).to(torch.bfloat16).cuda()
        imp, idx, sp, op = ops.PpnPlan(B, N, 100, 'cuda').run_embeds_bf16(s, o)
        torch.cuda.synchronize()
        assert float((imp - torch.matmul(s.float(), o.float().transpose(1, 2))).abs().max()) < 1e-5
    print('pair_topk<bf16> ok')
    st = torch.cuda.current_stream().cuda_stream
    for M, N, K in ((1500, 288, 256), (700, 256, 1024)):
        x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') * 0.05; b = torch.randn(N, device='cuda')
        wh = torch.empty((N, K), dtype=torch.bfloat16, device='cuda'); wl = torch.empty_like(wh); y = torch.empty(M, N, device='cuda')
        nat.check(lib.pn_split_bf16(w.data_ptr(), wh.data_ptr(), wl.data_ptr(), w.numel(), st), 'split16')
        nat.check(lib.pn_linear_tc_bf16x3(x.data_ptr(), wh.data_ptr(), wl.data_ptr(), b.data_ptr(), y.data_ptr(), N, M, N, K, st), 'b16')
        torch.cuda.synchronize()
        assert float((y - torch.addmm(b, x, w.t())).abs().max()) < 1e-3
    print('umma_gemm<W16> ok')
    from tests.util import oracle_small_head, product_small_head
    from oracle.make_golden import small_head_inputs
    head = product_small_head(oracle_small_head())
    mf, mems = small_head_inputs(2, (32, 48), 21)
    for _ in range(2):
        cls, msk = head.forward_from_memories(mf.cuda(), [m.cuda() for m in mems])
    torch.cuda.synchronize()
    assert bool(torch.isfinite(cls['rel']).all())
    print('head forward with PDL ok')
