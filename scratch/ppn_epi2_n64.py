import sys, statistics, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
import bench
from pairnet_b200 import _native as nat, ops
lib = nat.load()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
for N, Bm in ((64, 8192), (80, 6144), (100, 4096)):
    g = torch.Generator().manual_seed(1)
    s = F.normalize(torch.randn(Bm, N, 256, generator=g)).cuda(); o = F.normalize(torch.randn(Bm, N, 256, generator=g)).cuda()
    plan = ops.PpnPlan(Bm, N, 50, 'cuda')
    for opt in (1, 2, 1, 2):
        lib.pn_set_option(nat.PN_OPT_PPN_EPI2, opt)
        for _ in range(3): plan.run_embeds(s, o)
        ms = statistics.mean(bench.time_steps(lambda: plan.run_embeds(s, o), 10, flush, torch.cuda.current_stream()))
        by = 2 * N * 256 * 4 + N * N * 4 + 2 * 50 * 8
        print(f"N={N} fp32 groups-opt={opt}: {ms*1e3:.1f} us, {Bm*by/ms/1e6:.0f} GB/s", flush=True)
lib.pn_set_option(nat.PN_OPT_PPN_EPI2, 1)
