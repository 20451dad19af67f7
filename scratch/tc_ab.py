import torch, sys
sys.path.insert(0, '.')
from pairnet_b200 import _native as nat
lib = nat.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64*1024*1024, device='cuda')
def run(M,N,K, tag):
    x = torch.randn(M,K,device='cuda'); w = torch.randn(N,K,device='cuda')*0.05; b = torch.zeros(N,device='cuda'); y = torch.empty(M,N,device='cuda')
    xh,xl,wh,wl = torch.empty_like(x),torch.empty_like(x),torch.empty_like(w),torch.empty_like(w)
    lib.pn_split_tf32(x.data_ptr(),xh.data_ptr(),xl.data_ptr(),x.numel(),st); lib.pn_split_tf32(w.data_ptr(),wh.data_ptr(),wl.data_ptr(),w.numel(),st)
    def call(): nat.check(lib.pn_linear_tc_presplit(xh.data_ptr(),xl.data_ptr(),wh.data_ptr(),wl.data_ptr(),b.data_ptr(),y.data_ptr(),N,M,N,K,3,st),"tc")
    for _ in range(3): call()
    tot=0
    for _ in range(10):
        flush.add_(1); s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record(); call(); e.record(); torch.cuda.synchronize(); tot+=s.elapsed_time(e)
    print(f"{tag} M={M} N={N} K={K}: {tot/10*1e3:.1f} us  {2*M*N*K/(tot/10)/1e9:.0f} TF/s alg")
for wide, epi8 in ((0,0),(0,1),(1,1)):
    lib.pn_set_option(1, wide); lib.pn_set_option(2, epi8)
    for shp in ((33400,512,256),(43900,256,256),(43900,1024,256),(43900,256,1024),(43900,288,256)):
        run(*shp, f"wide={wide} epi8={epi8}")
