import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
dev='cuda'
def bench(M,N,K,n=50, distinct_weights=False):
    x = torch.randn(M,K,device=dev); b = torch.zeros(N,device=dev); y = torch.empty(M,N,device=dev)
    ws = [torch.randn(N,K,device=dev)*0.05 for _ in range(n if distinct_weights else 1)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        st = s.cuda_stream
        def call(i): 
            w = ws[i % len(ws)]
            nat.check(lib.pn_linear(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),None,y.data_ptr(),N,M,N,K,0,st),"l")
        for i in range(3): call(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n): call(i)
        g.replay(); torch.cuda.synchronize()
        e0=torch.cuda.Event(True); e1=torch.cuda.Event(True)
        e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
        t_warm = e0.elapsed_time(e1)/n*1e3
        flush = torch.empty(64*1024*1024, device=dev)
        flush.add_(1); torch.cuda.synchronize()
        e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
        t_cold = e0.elapsed_time(e1)/n*1e3
    return t_warm, t_cold
for shp in [(200,256,256),(200,2048,256),(200,512,256),(200,56,256)]:
    print(shp, "graph replay same weights: warm %.2f us/kernel, after L2 flush %.2f" % bench(*shp),
          "| 50 distinct weights: warm %.2f, after flush %.2f" % bench(*shp, distinct_weights=True))
