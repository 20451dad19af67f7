import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from pairnet_b200 import ops
B, N = 300, 100
g = torch.Generator().manual_seed(1)
s = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1).to(torch.bfloat16).cuda()
o = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1).to(torch.bfloat16).cuda()
imp, idx, sp, op = ops.PpnPlan(B, N, 100, 'cuda').run_embeds_bf16(s, o)
torch.cuda.synchronize()
print("ok", float((imp.double() - torch.matmul(s.double(), o.double().transpose(1, 2))).abs().max()))
