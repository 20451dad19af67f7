"""Relation Fusion stage: fused chain vs per-op kernels, CUDA-graph replayed, CUDA-event timed; + whole head."""
import sys, ctypes as C
import torch
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
from tests.test_gpu_chain import _build
lib = nat.load()
dev = 'cuda'

def time_graph(fn, iters=50):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000.0

for (N, R, B) in [(100, 100, 2), (100, 100, 8), (100, 100, 16), (200, 200, 2)]:
    o, p = _build(N, R)
    w = p.native_weights()
    pair = torch.randn((B, 2 * R, 256), device=dev)
    need = lib.pn_relation_fusion_workspace_bytes(B, R, 2 * R, 2048)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    rel = torch.zeros((B, R, 56), device=dev)
    def run():
        st = torch.cuda.current_stream().cuda_stream
        nat.check(lib.pn_relation_fusion_forward(C.byref(w.rel), pair.data_ptr(), rel.data_ptr(), None, B, 2 * R,
                                                 ws.data_ptr(), need, st), "rel")
    res = {}
    for fused in (2, 1, 0):
        lib.pn_set_option(nat.PN_OPT_FUSED_CHAIN, fused)
        res[fused] = (time_graph(run), lib.pn_last_launch_count())
    lib.pn_set_option(nat.PN_OPT_FUSED_CHAIN, 1)
    print(f"relation fusion N={N} R={R} B={B}: chain {res[2][0]:.1f} us ({res[2][1]} launches)  auto {res[1][0]:.1f} us ({res[1][1]})  per-op r01 {res[0][0]:.1f} us ({res[0][1]} launches)")

# whole head at bench shapes
from tests.test_gpu_head import _full_size_inputs
from tests.util import oracle_small_head, product_small_head
o = oracle_small_head(); p = product_small_head(o)
mf, mems = _full_size_inputs(2, 55)
mf = mf.cuda().contiguous(memory_format=torch.channels_last); mems = [m.cuda() for m in mems]
def head():
    p.forward_from_memories(mf, mems)
for fused in (2, 1, 0):
    lib.pn_set_option(nat.PN_OPT_FUSED_CHAIN, fused)
    t = time_graph(head, 20)
    print(f"head bs=2 full size, fused_chain={fused}: {t:.1f} us, {p.last_launch_count} launches")
lib.pn_set_option(nat.PN_OPT_FUSED_CHAIN, 1)
