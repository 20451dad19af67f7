"""Per-phase times of the fused decoder-chain kernel (globaltimer at every cluster barrier, cluster 0 / rank 0)."""
import sys, ctypes as C
import torch
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
from tests.test_gpu_chain import _build
lib = nat.load()
N, R, B = 100, int(sys.argv[1]) if len(sys.argv) > 1 else 100, 2
o, p = _build(N, R)
w = p.native_weights()
pair = torch.randn((B, 2 * R, 256), device='cuda')
need = lib.pn_relation_fusion_workspace_bytes(B, R, 2 * R, 2048)
ws = torch.empty(need, dtype=torch.uint8, device='cuda')
rel = torch.zeros((B, R, 56), device='cuda')
tb = torch.zeros(1024, dtype=torch.int64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def run():
    nat.check(lib.pn_relation_fusion_forward(C.byref(w.rel), pair.data_ptr(), rel.data_ptr(), None, B, 2 * R,
                                             ws.data_ptr(), need, st), "rel")
lib.pn_set_option(nat.PN_OPT_FUSED_CHAIN, 2)
for _ in range(3): run()
torch.cuda.synchronize()
lib.pn_debug_chain_timing(tb.data_ptr(), 1024)
run(); torch.cuda.synchronize()
lib.pn_debug_chain_timing(None, 0)
t = tb.cpu().tolist()
n = max(i for i, v in enumerate(t[:256]) if v) + 1
names = ["init"] + ["q", "xattn", "xout", "ln0", "qk", "v", "sattn", "sout", "ln1", "ffn1", "ffn2", "ln2"] * 6 + ["cls"]
print("total %.1f us over %d phases" % ((t[n - 1] - t[0]) / 1000.0, n - 1))
for i in range(1, min(n, 27)):
    step = i - 2
    extra = ""
    if step >= 0 and t[256 + step * 8]:
        tr = t[256 + step * 8: 256 + step * 8 + 4]
        extra = "  work(thread0) %.2f  fence %.2f  arrive %.2f  wait %.2f" % ((tr[0] - t[i - 1]) / 1e3, (tr[1] - tr[0]) / 1e3, (tr[2] - tr[1]) / 1e3, (tr[3] - tr[2]) / 1e3)
    print("%2d %-6s %7.2f us%s" % (i, names[i - 1] if i - 1 < len(names) else "?", (t[i] - t[i - 1]) / 1000.0, extra))
