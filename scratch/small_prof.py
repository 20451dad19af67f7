"""One query-side linear for ncu.  argv: M N K"""
import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (200, 256, 256)
x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') * 0.05; b = torch.zeros(N, device='cuda'); y = torch.empty(M, N, device='cuda')
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
for _ in range(6):
    nat.check(lib.pn_linear(x.data_ptr(), K, w.data_ptr(), b.data_ptr(), None, y.data_ptr(), N, M, N, K, 1, st), "l")
torch.cuda.synchronize()
