import torch, sys, os
sys.path.insert(0, '.')
import torch.nn.functional as F
from bench import build_model, synthetic_images
dev = torch.device('cuda')
model = build_model(dev); head = model.bbox_head; pd = head.pixel_decoder
img = synthetic_images(2, 1).to(dev)
def T(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n
with torch.no_grad():
    feats = model.extract_feat(img)
    print("backbone", T(lambda: model.extract_feat(img)))
    print("pixel decoder total", T(lambda: pd(feats)))
    # pieces
    def inconv():
        return [pd.input_convs[i](feats[3 - i]).flatten(2).transpose(1, 2) for i in range(3)]
    print("  input convs+GN+flatten", T(inconv))
    xs = inconv(); shapes=[tuple(feats[3-i].shape[-2:]) for i in range(3)]
    pos_l, ref, norm = pd._geometry(shapes, dev, torch.float32)
    pos = torch.cat([p + pd.level_encoding.weight[i][None, :] for i, p in pos_l], 0)[None]
    x = torch.cat(xs, 1).contiguous()
    print("  cat+pos", T(lambda: (torch.cat(inconv.__call__(),1), torch.cat([p + pd.level_encoding.weight[i][None, :] for i, p in pos_l], 0))) )
    print("  encoder native", T(lambda: pd._native_encoder(x, pos[0].contiguous(), shapes)))
    mem = pd._native_encoder(x, pos[0].contiguous(), shapes)
    def fpn():
        m = mem.transpose(1,2); start=0; outs=[]
        for h,w in shapes:
            outs.append(m[:, :, start:start+h*w].reshape(2,-1,h,w)); start+=h*w
        cur = pd.lateral_convs[0](feats[0])
        y = cur + F.interpolate(outs[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
        o = pd.output_convs[0](y)
        return pd.mask_feature(o), outs
    print("  fpn+mask_feature", T(fpn))
    cur = pd.lateral_convs[0](feats[0])
    print("    lateral conv+GN", T(lambda: pd.lateral_convs[0](feats[0])))
    print("    lateral conv only", T(lambda: pd.lateral_convs[0].conv(feats[0])))
    c = pd.lateral_convs[0].conv(feats[0])
    print("    GN only", T(lambda: pd.lateral_convs[0].gn(c)), c.is_contiguous(), c.is_contiguous(memory_format=torch.channels_last))
    print("    output conv3x3 only", T(lambda: pd.output_convs[0].conv(cur)))
    print("    mask_feature conv", T(lambda: pd.mask_feature(cur)))
    mf, mems = pd(feats)
    print("mf contiguous", mf.is_contiguous(), [m.is_contiguous() for m in mems])
    print("head", T(lambda: head.forward_from_memories(mf, mems)))
    print("head no seg", T(lambda: head.forward_from_memories(mf, mems, materialize_seg=False)))
