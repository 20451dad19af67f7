import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
dev='cuda'
def run(M,N,K,relu=0,splits=1):
    x = torch.randn(M,K,device=dev); w = torch.randn(N,K,device=dev)*0.05; b = torch.randn(N,device=dev); r = torch.randn(M,N,device=dev)
    ref = (x.double()@w.double().T + b.double())
    if relu: ref = ref.relu()
    ref = ref + r.double()
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for opt in (0,1,11,12,14,22,24):
        lib.pn_set_option(8,opt)
        y = torch.empty(M,N,device=dev)
        nat.check(lib.pn_linear(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),r.data_ptr(),y.data_ptr(),N,M,N,K,relu,st),"l")
        torch.cuda.synchronize()
        err = ((y.double()-ref).abs().max()/ref.abs().max()).item()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(50):
                    nat.check(lib.pn_linear(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),r.data_ptr(),y.data_ptr(),N,M,N,K,relu,s.cuda_stream),"l")
            g.replay(); torch.cuda.synchronize()
            e0=torch.cuda.Event(True); e1=torch.cuda.Event(True)
            e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
        res.append("%d: %.1e %.2fus" % (opt, err, e0.elapsed_time(e1)/50*1e3))
    lib.pn_set_option(8,1)
    print((M,N,K,relu), " | ".join(res))
for shp in [(200,256,256,0),(200,2048,256,1),(200,512,256,0),(200,768,256,0),(200,56,256,0),(400,256,256,0),(400,2048,256,1),(800,768,256,0),(1000,2048,256,0)]:
    run(*shp)
