"""ConvTiny: tcgen05 conv2 path vs FFMA path, per-kernel CUDA-event timing."""
import sys, statistics, torch
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat, ops
from oracle.head import OConvTiny
lib = nat.load()
torch.manual_seed(1)
m = OConvTiny(mid_channels=64).cuda()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
for B, N in ((2, 100), (2, 200), (2, 400), (32, 100)):
    x = torch.tanh(torch.randn(B, N, N, device='cuda'))
    for tc in (1, 0):
        lib.pn_set_option(nat.PN_OPT_CONV_TC, tc)
        for _ in range(2): ops.conv_tiny(x, m)
        ts = []
        for _ in range(5):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.conv_tiny(x, m); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = statistics.mean(ts)
        fl = B * 413952.0 * N * N
        print(f"B={B} N={N} tc={tc}: {t*1e3:.1f} us  {fl/t/1e9:.1f} TFLOP/s (whole ConvTiny incl. workspace alloc in the wrapper)")
lib.pn_set_option(nat.PN_OPT_CONV_TC, 1)
