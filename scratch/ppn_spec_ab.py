import sys, statistics, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
import bench
from pairnet_b200 import _native as nat, ops
lib = nat.load()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
for N, Bm in ((100, 4096), (200, 2048)):
    g = torch.Generator().manual_seed(1234)
    s = F.normalize(torch.randn(Bm, N, 256, generator=g)).cuda(); o = F.normalize(torch.randn(Bm, N, 256, generator=g)).cuda()
    sb, ob = s.to(torch.bfloat16), o.to(torch.bfloat16)
    plan = ops.PpnPlan(Bm, N, 100, 'cuda')
    for opt in (0, 1, 0, 1):
        lib.pn_set_option(nat.PN_OPT_PPN_SPECULATE, opt)
        for _ in range(3): plan.run_embeds(s, o)
        ms = statistics.mean(bench.time_steps(lambda: plan.run_embeds(s, o), 10, flush, torch.cuda.current_stream()))
        for _ in range(3): plan.run_embeds_bf16(sb, ob)
        ms16 = statistics.mean(bench.time_steps(lambda: plan.run_embeds_bf16(sb, ob), 10, flush, torch.cuda.current_stream()))
        redo = torch.frombuffer(plan.ws.cpu().numpy().tobytes(), dtype=torch.int32) if False else None
        print(f"N={N} speculate={opt}: fp32 {ms*1e3:.1f} us, bf16 {ms16*1e3:.1f} us", flush=True)
lib.pn_set_option(nat.PN_OPT_PPN_SPECULATE, 1)
