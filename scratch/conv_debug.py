import sys, torch
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat, ops
from oracle.head import OConvTiny
torch.manual_seed(1)
torch.backends.cudnn.allow_tf32 = False
m = OConvTiny(mid_channels=64).cuda()
B, N = 1, 40
x = torch.tanh(torch.randn(B, N, N, device='cuda'))
c2 = m.conv_layers[1][0]
w_orig = c2.weight.detach().clone()
def probe(name):
    with torch.no_grad():
        ref = m(x)
    got = ops.conv_tiny(x, m)
    e = float((got - ref).abs().max()) / float(ref.abs().max())
    print(f"{name}: rel err {e:.3e}")
with torch.no_grad():
    c2.bias.zero_()
    for (ky, kx) in [(3, 3), (3, 0), (3, 1), (3, 6), (0, 3), (6, 3), (0, 0), (6, 6), (2, 5)]:
        c2.weight.zero_()
        for c in range(64): c2.weight[c, c, ky, kx] = 1.0
        probe(f"delta tap ky={ky} kx={kx} identity channels")
    c2.weight.zero_()
    c2.weight[:, :, 3, 3] = w_orig[:, :, 3, 3]
    probe("centre tap only, full channel mixing")
    c2.weight.zero_()
    c2.weight[:, :32, 3, 3] = w_orig[:, :32, 3, 3]
    probe("centre tap, input channels 0-31 only")
    c2.weight.zero_()
    c2.weight[:, 32:, 3, 3] = w_orig[:, 32:, 3, 3]
    probe("centre tap, input channels 32-63 only")
    c2.weight.copy_(w_orig)
    probe("full weights")
