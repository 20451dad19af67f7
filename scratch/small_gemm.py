import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import ops, _native as nat
lib = nat.load()
st = torch.cuda.current_stream().cuda_stream
def bench(M,N,K, flush=False, n=50):
    x = torch.randn(M,K,device='cuda'); w = torch.randn(N,K,device='cuda')*0.05; b = torch.zeros(N,device='cuda'); y = torch.empty(M,N,device='cuda')
    big = torch.empty(64*1024*1024, device='cuda')
    def call(): nat.check(lib.pn_linear(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),None,y.data_ptr(),N,M,N,K,0,st),"l")
    for _ in range(5): call()
    torch.cuda.synchronize()
    if not flush:
        s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record()
        for _ in range(n): call()
        e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n*1e3
    tot=0
    for _ in range(n):
        big.add_(1); s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record(); call(); e.record(); torch.cuda.synchronize(); tot+=s.elapsed_time(e)
    return tot/n*1e3
for shp in [(200,256,256),(200,2048,256),(200,512,256),(400,256,256),(200,56,256)]:
    print(shp, "warm back-to-back %.2f us" % bench(*shp), " cold (L2 flushed) %.2f us" % bench(*shp, flush=True))
# empty-ish kernel reference: layernorm on 200 rows
x = torch.randn(200,256,device='cuda'); g=torch.ones(256,device='cuda'); bb=torch.zeros(256,device='cuda')
for _ in range(5): ops.add_layernorm(x,None,g,bb)
torch.cuda.synchronize(); s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record()
for _ in range(50): ops.add_layernorm(x,None,g,bb)
e.record(); torch.cuda.synchronize(); print("layernorm warm (incl. torch.empty) %.2f us" % (s.elapsed_time(e)/50*1e3))
