import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64*1024*1024, device='cuda')
def run(M,N,K, raw):
    lib.pn_set_option(5, raw)
    x = torch.randn(M,K,device='cuda'); w = torch.randn(N,K,device='cuda')*0.05; b = torch.randn(N,device='cuda'); y = torch.empty(M,N,device='cuda')
    need = lib.pn_linear_tc_workspace_bytes(M,N,K); ws = torch.empty(need,dtype=torch.uint8,device='cuda')
    def call(): nat.check(lib.pn_linear_tc(x.data_ptr(),K,w.data_ptr(),b.data_ptr(),y.data_ptr(),N,M,N,K,3,ws.data_ptr(),need,st),"tc")
    for _ in range(3): call()
    tot=0
    for _ in range(10):
        flush.add_(1); s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record(); call(); e.record(); torch.cuda.synchronize(); tot+=s.elapsed_time(e)
    ref = (x.double()@w.double().t()+b.double()).float()
    err = float((y-ref).abs().max()/ref.abs().max())
    print(f"raw={raw} M={M} N={N} K={K}: {tot/10*1e3:.1f} us incl. splits  {2*M*N*K/(tot/10)/1e9:.0f} TF/s alg  err {err:.2e}")
for epi8 in (0, 1):
    lib.pn_set_option(2, epi8)
    print("epi8", epi8)
    for shp in ((33400,512,256),(43900,256,256),(43900,1024,256),(43900,256,1024),(43900,288,256)):
        run(*shp, 1)
