"""fa_umma_kernel (+ combine) time at the three memory levels, CUDA events, L2 flushed; and the head graph time."""
import sys, statistics, torch
sys.path.insert(0, '.')
import bench
from pairnet_b200 import _native as nat, ops
from pairnet_b200.detector import GraphedForward
dev = torch.device("cuda", 0)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for Nk in (1050, 4200, 16700, 200):
    B, Nq = 2, 100
    q = torch.randn(B, Nq, 256, device=dev) * 0.3; k = torch.randn(B, Nk, 256, device=dev) * 0.3; v = torch.randn(B, Nk, 256, device=dev)
    E = torch.randn(B, Nq, 256, device=dev); Fl = torch.randn(B, 256, (Nk + 63) // 64 * 64, device=dev)
    bits, rowany = ops.attn_mask_bits(E, Fl, Nk)
    run = lambda: ops.mha_core_tc(q, k, v, bits, rowany)
    for _ in range(3): run()
    print(f"Nk={Nk}: mha_core_tc (split + fa + combine) {1e3 * statistics.mean(bench.time_steps(run, 20, flush, torch.cuda.current_stream())):.1f} us")
model = bench.build_model(dev)
imgs = bench.synthetic_images(2, 1).to(dev)
with torch.no_grad():
    feats = model.extract_feat(imgs)
    mf, mems = model.bbox_head.pixel_decoder(feats)
    head = model.bbox_head
    gh = GraphedForward(lambda x: head.forward_from_memories(mf, mems), imgs)
print(f"head graph {statistics.mean(bench.time_steps(lambda: gh(), 40, flush, torch.cuda.current_stream())):.4f} ms")
