"""PPN 5a: fused pair-matrix + top-k kernel vs the two-kernel path, CUDA-event timed with L2 flush."""
import sys, statistics, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat, ops
lib = nat.load()
dev = 'cuda'
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def time_it(fn, n=10):
    ts = []
    for _ in range(n):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.mean(ts), min(ts)
for N, Bm in ((100, 4096), (200, 2048), (400, 1024), (100, 2)):
    g = torch.Generator().manual_seed(1234)
    s = F.normalize(torch.randn(Bm, N, 256, generator=g)).to(dev)
    o = F.normalize(torch.randn(Bm, N, 256, generator=g)).to(dev)
    plan = ops.PpnPlan(Bm, N, 100, dev)
    by = Bm * (2 * N * 256 * 4 + N * N * 4 + 2 * 100 * 8)
    for fused in (1, 0):
        lib.pn_set_option(nat.PN_OPT_PPN_FUSED_TOPK, fused)
        for _ in range(3): plan.run_embeds(s, o)
        m, mn = time_it(lambda: plan.run_embeds(s, o))
        print(f"N={N} B={Bm} fused={fused}: mean {m*1e3:.1f} us min {mn*1e3:.1f} us  {by/m/1e6:.0f} GB/s ({by/m/1e6/6535.7:.3f} of HBM peak)")
    lib.pn_set_option(nat.PN_OPT_PPN_FUSED_TOPK, 1)
# L2-resident variant (no flush, inputs + outputs of 296 images = 73 MB): separates DRAM behaviour from the SM pipeline
for N, Bm in ((100, 296), (100, 592)):
    g = torch.Generator().manual_seed(1234)
    s = F.normalize(torch.randn(Bm, N, 256, generator=g)).to(dev)
    o = F.normalize(torch.randn(Bm, N, 256, generator=g)).to(dev)
    plan = ops.PpnPlan(Bm, N, 100, dev)
    for _ in range(5): plan.run_embeds(s, o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): plan.run_embeds(s, o)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    print(f"L2-resident N={N} B={Bm}: {t*1e3:.1f} us per call = {t*1e3/(Bm/148):.2f} us per image per SM")
