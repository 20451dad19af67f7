"""Head graph time with PN_OPT_UMMA_EPI8 (8 epilogue warps in the 3xTF32 raw-A GEMM: mask bits, K projections) on/off."""
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
for opt in (0, 1, 0, 1):
    lib.pn_set_option(2, opt)
    with torch.no_grad():
        gh = GraphedForward(lambda x: head.forward_from_memories(mf, mems), imgs)
    th = statistics.mean(bench.time_steps(lambda: gh(), 40, flush, torch.cuda.current_stream()))
    print(f"PN_OPT_UMMA_EPI8={opt}: head {th:.4f} ms", flush=True)
    del gh
lib.pn_set_option(2, 0)
