import sys, os, torch, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
bb = model.backbone
x = bench.synthetic_images(2, 1).to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
res = {}
with torch.no_grad():
    for mode in (False, True):
        type(bb).fuse_epilogues = mode
        for _ in range(3): outs = bb(x)
        ts = bench.time_steps(lambda: bb(x), 10, flush, torch.cuda.current_stream())
        res[mode] = [o.clone() for o in outs]
        print("fuse_epilogues", mode, f"{statistics.mean(ts):.3f} ms", [tuple(o.shape) for o in outs], outs[0].is_contiguous(memory_format=torch.channels_last))
    for a, b in zip(res[False], res[True]):
        print("rel diff", float((a - b).abs().max() / a.abs().max()))
