import torch, sys, collections
sys.path.insert(0, '.')
from bench import build_model, synthetic_images
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
model = build_model(dev); head = model.bbox_head
img = synthetic_images(2, 1).to(dev)
with torch.no_grad():
    feats = model.extract_feat(img); mf, mems = head.pixel_decoder(feats)
    mems = [m.contiguous() for m in mems]
    for _ in range(3): head.forward_from_memories(mf, mems)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): head.forward_from_memories(mf, mems)
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[e.name[:80]]; a[0] += 1; a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print("head total kernel time per forward: %.1f us" % (tot / 5))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
    print("%9.1f us/fwd %5.1f/fwd %8.2f us each  %s" % (t / 5, n / 5, t / n, k))
