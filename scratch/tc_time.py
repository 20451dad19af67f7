import torch, time, sys
sys.path.insert(0, '.')
from pairnet_b200 import ops, _native as nat
lib = nat.load()
M,N,K = 33400, 512, 256
x = torch.randn(M,K,device='cuda'); w = torch.randn(N,K,device='cuda')*0.05; b = torch.zeros(N,device='cuda')
need = lib.pn_linear_tc_workspace_bytes(M,N,K); ws = torch.empty(need,dtype=torch.uint8,device='cuda'); y = torch.empty(M,N,device='cuda')
st = torch.cuda.current_stream().cuda_stream
for passes in (3,1):
    for _ in range(3): nat.check(lib.pn_linear_tc(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),y.data_ptr(),N,M,N,K,passes,ws.data_ptr(),need,st),"tc")
    torch.cuda.synchronize(); s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record()
    for _ in range(20): nat.check(lib.pn_linear_tc(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),y.data_ptr(),N,M,N,K,passes,ws.data_ptr(),need,st),"tc")
    e.record(); torch.cuda.synchronize(); ms=s.elapsed_time(e)/20
    print(f"passes={passes}: {ms*1e3:.1f} us incl. split kernels -> {2*M*N*K/ms/1e9:.1f} TFLOP/s (algorithmic)")
ref = (x.double()@w.double().t()).float()
print("err", float((y-ref).abs().max()/ref.abs().max()))
y2 = ops.linear(x,w,b); torch.cuda.synchronize()
s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record()
for _ in range(20): ops.linear(x,w,b)
e.record(); torch.cuda.synchronize(); print("simt us", s.elapsed_time(e)/20*1e3)
