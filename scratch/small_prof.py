import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
M,N,K=200,256,256
x = torch.randn(M,K,device='cuda'); w = torch.randn(N,K,device='cuda')*0.05; b = torch.zeros(N,device='cuda'); y = torch.empty(M,N,device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(6): nat.check(lib.pn_linear(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),None,y.data_ptr(),N,M,N,K,0,st),"l")
torch.cuda.synchronize()
