import sys, torch
sys.path.insert(0, '.')
from pairnet_b200 import ops
from oracle.head import OConvTiny
torch.manual_seed(1)
m = OConvTiny(mid_channels=64).cuda()
x = torch.tanh(torch.randn(2, 100, 100, device='cuda'))
for _ in range(3): ops.conv_tiny(x, m)
torch.cuda.synchronize()
