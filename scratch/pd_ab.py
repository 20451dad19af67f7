import torch, sys
sys.path.insert(0, '.')
from bench import build_model, synthetic_images
from pairnet_b200 import _native as nat
lib = nat.load()
dev = torch.device('cuda')
model = build_model(dev); head = model.bbox_head; pd = head.pixel_decoder
img = synthetic_images(2, 1).to(dev)
flush = torch.empty(64*1024*1024, device=dev)
def T(fn, n=10):
    for _ in range(3): fn()
    tot=0
    for _ in range(n):
        flush.add_(1); s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record(); fn(); e.record(); torch.cuda.synchronize(); tot+=s.elapsed_time(e)
    return tot/n
with torch.no_grad():
    feats = model.extract_feat(img)
    mf, mems = pd(feats)
    for epi8 in (0,1,0,1):
        lib.pn_set_option(2, epi8)
        print("epi8", epi8, "pixel decoder %.3f ms" % T(lambda: pd(feats)), "head %.3f ms" % T(lambda: head.forward_from_memories(mf, mems)))
