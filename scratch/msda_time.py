"""Pixel-decoder time in the replayed graph + msda_sample kernel time (CUDA events over the standalone entry)."""
import sys, statistics, torch, ctypes as C
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
with torch.no_grad():
    feats = model.extract_feat(imgs)
    head = model.bbox_head
    gp = GraphedForward(lambda x: head.pixel_decoder(feats), imgs)
    gf = GraphedForward(model.forward_dummy, imgs)
tp = statistics.mean(bench.time_steps(lambda: gp(), 30, flush, torch.cuda.current_stream()))
tf = statistics.mean(bench.time_steps(lambda: gf(), 30, flush, torch.cuda.current_stream()))
print(f"pixel decoder {tp:.4f} ms, whole forward {tf:.4f} ms")
# standalone sampler
B, hs, ws = 2, [25, 50, 100], [42, 84, 167]
nq = sum(h * w for h, w in zip(hs, ws))
value = torch.randn(B, nq, 256, device=dev)
ol = torch.randn(B * nq, 8 * 3 * 4 * 3, device=dev)
out = torch.empty(B * nq, 256, device=dev)
h = (C.c_int * 3)(*hs); w = (C.c_int * 3)(*ws)
st = torch.cuda.current_stream().cuda_stream
run = lambda: nat.check(lib.pn_msda_sample(value.data_ptr(), ol.data_ptr(), out.data_ptr(), h, w, 3, 4, B, st), "msda")
for _ in range(3): run()
print(f"msda_sample_kernel {1e3 * statistics.mean(bench.time_steps(run, 20, flush, torch.cuda.current_stream())):.1f} us")
