"""Two batches in flight: two CUDA graphs (two model replicas, own workspaces) replayed on two streams vs one graph."""
import sys, time, torch
sys.path.insert(0, '.')
import bench
from pairnet_b200.detector import GraphedForward
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
models = [bench.build_model(dev) for _ in range(2)]
imgs = [bench.synthetic_images(2, s).to(dev) for s in (1, 2)]
with torch.no_grad():
    runners = [GraphedForward(m.forward_dummy, x) for m, x in zip(models, imgs)]
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
def run(K, two):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for i in range(K):
        j = i % 2 if two else 0
        with torch.cuda.stream(streams[j]):
            runners[j]()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
for two in (False, True, False, True):
    run(10, two)
    ms = run(60, two)
    print(f"two_in_flight={two}: {ms:.3f} ms per bs=2 step -> {2e3 / ms:.1f} img/s", flush=True)
