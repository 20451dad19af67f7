"""Head graph time vs the skinny-GEMM tile used for the wide (FFN) problems (GPU box)."""
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
with torch.no_grad():
    feats = model.extract_feat(imgs)
    mf, mems = model.bbox_head.pixel_decoder(feats)
head = model.bbox_head
for opt in (1, 112, 114, 122, 124, 1, 124, 122):
    lib.pn_set_option(nat.PN_OPT_SKINNY, opt)
    with torch.no_grad():
        gh = GraphedForward(lambda x: head.forward_from_memories(mf, mems), imgs)
    th = statistics.mean(bench.time_steps(lambda: gh(), 40, flush, torch.cuda.current_stream()))
    print(f"PN_OPT_SKINNY={opt}: head {th:.4f} ms", flush=True)
    del gh
lib.pn_set_option(nat.PN_OPT_SKINNY, 1)
