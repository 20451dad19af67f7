import sys, collections, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from pairnet_b200 import ops
from torch.profiler import profile, ProfilerActivity
dev = 'cuda'
for N, Bm in ((100, 4096), (200, 2048), (400, 1024)):
    g = torch.Generator(device="cpu").manual_seed(1234)
    s = F.normalize(torch.randn(Bm, N, 256, generator=g)).to(dev); o = F.normalize(torch.randn(Bm, N, 256, generator=g)).to(dev)
    plan = ops.PpnPlan(Bm, N, 100, dev)
    for _ in range(3): plan.run_embeds(s, o)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        plan.run_embeds(s, o); torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    for e in evs:
        a = agg[e.name[:50]]; a[0] += 1; a[1] += e.device_time
    span = evs[-1].time_range.end - evs[0].time_range.start
    print(N, Bm, "span us", round(span, 1), {k: (v[0], round(v[1], 1)) for k, v in agg.items()})
