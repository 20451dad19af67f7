"""In-graph timeline of the HEAD alone (CUPTI): per-kernel durations, union busy time, idle gaps, critical-chain view."""
import sys, json, collections
import torch
sys.path.insert(0, '.')
from bench import build_model, synthetic_images
from pairnet_b200.detector import GraphedForward
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
from pairnet_b200 import _native as nat
if len(sys.argv) > 1:
    nat.load().pn_set_option(nat.PN_OPT_PDL, int(sys.argv[1]))
    print("PN_OPT_PDL =", sys.argv[1])
model = build_model(dev); head = model.bbox_head
img = synthetic_images(2, 1).to(dev)
with torch.no_grad():
    feats = model.extract_feat(img)
    mf, mems = head.pixel_decoder(feats)
    runner = GraphedForward(lambda x: head.forward_from_memories(mf, mems), img, warmup=2)
    for _ in range(3): runner()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        runner(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs))
t0, t1 = rows[0][0], max(r[1] for r in rows)
print("head span %.1f us, %d kernels, sum of durations %.1f us" % (t1 - t0, len(rows), sum(e - s for s, e, _ in rows)))
# union busy time and gaps
busy, cur_s, cur_e, gaps = 0.0, rows[0][0], rows[0][1], []
for s, e, n in rows[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; gaps.append((s - cur_e, n[:50])); cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("union busy %.1f us, idle %.1f us in %d gaps (mean %.2f us)" % (busy, (t1 - t0) - busy, len(gaps), sum(g for g, _ in gaps) / max(1, len(gaps))))
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in rows:
    a = agg[n[:70]]; a[0] += 1; a[1] += e - s
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    print("%8.1f us %4d x %7.2f us  %s" % (t, n, t / n, k))
json.dump([dict(s=s - t0, e=e - t0, n=n[:60]) for s, e, n in rows], open('gpurun_out/trace_head.json', 'w'))
