"""A/B of PN_OPT_ENC_BF16X3 on the graph-replayed pixel decoder / whole forward (GPU box)."""
import sys, statistics, torch
sys.path.insert(0, '.')
import bench
from pairnet_b200 import _native as nat
from pairnet_b200.detector import GraphedForward
dev = torch.device("cuda", 0)
lib = nat.load()
model = bench.build_model(dev)
imgs = bench.synthetic_images(2, 1).to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
torch.backends.cudnn.benchmark = True
ref = None
for mode in (0, 1, 0, 1):
    lib.pn_set_option(nat.PN_OPT_ENC_BF16X3, mode)
    with torch.no_grad():
        feats = model.extract_feat(imgs)
        head = model.bbox_head
        mf, mems = head.pixel_decoder(feats)
        if ref is None:
            ref = [m.clone() for m in mems]
        else:
            print("  max rel diff of the memories vs first mode:", max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(mems, ref)))
        gp = GraphedForward(lambda x: head.pixel_decoder(feats), imgs)
        gf = GraphedForward(model.forward_dummy, imgs)
    tp = statistics.mean(bench.time_steps(lambda: gp(), 30, flush, torch.cuda.current_stream()))
    tf = statistics.mean(bench.time_steps(lambda: gf(), 30, flush, torch.cuda.current_stream()))
    print(f"ENC_BF16X3={mode}: pixel decoder {tp:.4f} ms, whole forward {tf:.4f} ms", flush=True)
    del gf, gp
