import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
M,N,K = 33400, 512, 256
x = torch.randn(M,K,device='cuda'); w = torch.randn(N,K,device='cuda')*0.05; b = torch.zeros(N,device='cuda')
need = lib.pn_linear_tc_workspace_bytes(M,N,K); ws = torch.empty(need,dtype=torch.uint8,device='cuda'); y = torch.empty(M,N,device='cuda')
st = torch.cuda.current_stream().cuda_stream
lib.pn_set_option(5, 1)
for _ in range(3): nat.check(lib.pn_linear_tc(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),y.data_ptr(),N,M,N,K,3,ws.data_ptr(),need,st),"tc")
torch.cuda.synchronize()
