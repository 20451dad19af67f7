"""One-kernel driver for ncu: PPN 5a fused kernel.  argv: N B [iters] [bf16]"""
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from pairnet_b200 import ops
N, B = int(sys.argv[1]), int(sys.argv[2])
it = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bf16 = len(sys.argv) > 4 and sys.argv[4] == "bf16"
g = torch.Generator().manual_seed(1234)
s = F.normalize(torch.randn(B, N, 256, generator=g)).cuda()
o = F.normalize(torch.randn(B, N, 256, generator=g)).cuda()
if bf16:
    s, o = s.to(torch.bfloat16), o.to(torch.bfloat16)
plan = ops.PpnPlan(B, N, 100, 'cuda')
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
for _ in range(it):
    flush.add_(1.0)
    (plan.run_embeds_bf16 if bf16 else plan.run_embeds)(s, o)
torch.cuda.synchronize()
