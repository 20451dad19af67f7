import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
B,H,W = 2,200,334
x = torch.randn(B,256,H,W,device='cuda').contiguous(memory_format=torch.channels_last)
w = torch.randn(256,256,device='cuda')*0.05; b = torch.randn(256,device='cuda')
need = lib.pn_conv1x1_nhwc_to_nchw_workspace_bytes(256); ws = torch.empty(need,dtype=torch.uint8,device='cuda')
y = torch.empty(B,256,H,W,device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    nat.check(lib.pn_conv1x1_nhwc_to_nchw(x.data_ptr(),w.data_ptr(),b.data_ptr(),y.data_ptr(),B,H*W,256,ws.data_ptr(),need,st),"c")
torch.cuda.synchronize()
e0=torch.cuda.Event(True); e1=torch.cuda.Event(True); e0.record()
for _ in range(10):
    nat.check(lib.pn_conv1x1_nhwc_to_nchw(x.data_ptr(),w.data_ptr(),b.data_ptr(),y.data_ptr(),B,H*W,256,ws.data_ptr(),need,st),"c")
e1.record(); torch.cuda.synchronize(); print("conv1x1 us", e0.elapsed_time(e1)*100)
