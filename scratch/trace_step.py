"""Timeline of graph-replayed steps (torch profiler / CUPTI): per-kernel in-situ durations, stream overlap, idle gaps."""
import sys, json, collections
import torch
sys.path.insert(0, '.')
from bench import build_model, synthetic_images
from pairnet_b200.detector import GraphedForward
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
model = build_model(dev); head = model.bbox_head
img = synthetic_images(2, 1).to(dev)
def forward(x):
    cls, msk = model.forward_dummy(x)
    return cls, msk
with torch.no_grad():
    forward(img); torch.cuda.synchronize()
    runner = GraphedForward(forward, img, warmup=2)
    for _ in range(3): runner()
    torch.cuda.synchronize()
    NREP = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(NREP):
            runner()
            torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs))
# split into replays by big gaps
out = [dict(s=s, e=e, n=n[:90]) for s, e, n in rows]
json.dump(out, open('gpurun_out/trace_step.json', 'w'))
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in rows:
    a = agg[n[:80]]; a[0] += 1; a[1] += (e - s)
tot = sum(v[1] for v in agg.values())
print("sum of kernel time per replay: %.1f us" % (tot / NREP))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print("%9.1f us/step %6.1f/step %8.2f us each  %s" % (t / NREP, n / NREP, t / n, k))
