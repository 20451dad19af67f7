"""A/B of PN_OPT_PDL and of the native 1x1 convolutions on the graph-replayed head / whole forward (GPU box)."""
import sys, statistics, torch
sys.path.insert(0, '.')
import bench
from pairnet_b200 import _native as nat
from pairnet_b200.detector import GraphedForward
from pairnet_b200.upstream.pixel_decoder import ConvModule
dev = torch.device("cuda", 0)
lib = nat.load()
model = bench.build_model(dev)
imgs = bench.synthetic_images(2, 1).to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
torch.backends.cudnn.benchmark = True
for pdl, c11 in ((0, True), (1, True), (1, False), (0, False), (1, True)):
    lib.pn_set_option(nat.PN_OPT_PDL, pdl)
    ConvModule.native_conv1x1 = c11
    with torch.no_grad():
        feats = model.extract_feat(imgs)
        mf, mems = model.bbox_head.pixel_decoder(feats)
        head = model.bbox_head
        gh = GraphedForward(lambda x: head.forward_from_memories(mf, mems), imgs)
        gp = GraphedForward(lambda x: head.pixel_decoder(feats), imgs)
        gf = GraphedForward(model.forward_dummy, imgs)
    th = statistics.mean(bench.time_steps(lambda: gh(), 30, flush, torch.cuda.current_stream()))
    tp = statistics.mean(bench.time_steps(lambda: gp(), 30, flush, torch.cuda.current_stream()))
    tf = statistics.mean(bench.time_steps(lambda: gf(), 30, flush, torch.cuda.current_stream()))
    print(f"PDL={pdl} native1x1={c11}: head {th:.4f} ms, pixel decoder {tp:.4f} ms, whole forward {tf:.4f} ms", flush=True)
    del gh, gf, gp
