"""Thin torch-tensor wrappers over the stage-level C-ABI entry points (``include/pairnet_b200.h``).
Used by the stage-wise parity tests and the PPN micro-benchmark; ``CrossHead2.forward`` itself makes
a single ``pn_head_forward`` call.  Everything here requires CUDA tensors; nothing falls back."""
import ctypes as C

import torch

from . import _native as nat


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _f32(t):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise nat.NativeError("pairnet_b200 ops need fp32 CUDA tensors (no CPU path)")
    return t.contiguous()


def sine_posenc(h, w, device="cuda"):
    out = torch.empty((h * w, 256), dtype=torch.float32, device=device)
    nat.check(nat.load().pn_sine_posenc(out.data_ptr(), h, w, _stream(out)), "pn_sine_posenc")
    return out


def level_prep(mem, level_embed, pos):
    mem, level_embed, pos = _f32(mem), _f32(level_embed), _f32(pos)
    B, Cc, h, w = mem.shape
    x = torch.empty((B, h * w, Cc), dtype=torch.float32, device=mem.device)
    xp = torch.empty_like(x)
    nat.check(nat.load().pn_level_prep(mem.data_ptr(), level_embed.data_ptr(), pos.data_ptr(), x.data_ptr(),
                                       xp.data_ptr(), B, h * w, _stream(mem)), "pn_level_prep")
    return x, xp


def mask_feature_resize(F_, h, w):
    F_ = _f32(F_)
    B, Cc, H, W = F_.shape
    ldo = (h * w + 63) // 64 * 64
    out = torch.empty((B, Cc, ldo), dtype=torch.float32, device=F_.device)
    nat.check(nat.load().pn_mask_feature_resize(F_.data_ptr(), out.data_ptr(), B, H, W, h, w, ldo, _stream(F_)),
              "pn_mask_feature_resize")
    return out


def attn_mask_bits(E, Fl, hw):
    E, Fl = _f32(E), _f32(Fl)
    B, N, _ = E.shape
    ldf = Fl.shape[2]
    bits = torch.zeros((B, N, ldf // 32), dtype=torch.int32, device=E.device)
    rowany = torch.zeros((B * N,), dtype=torch.int32, device=E.device)
    nat.check(nat.load().pn_attn_mask_bits(E.data_ptr(), Fl.data_ptr(), bits.data_ptr(), rowany.data_ptr(), B, N, hw,
                                           ldf, _stream(E)), "pn_attn_mask_bits")
    return bits, rowany


def unpack_bits(bits, hw):
    """[.., words] int32 -> [.., hw] bool (True = blocked)."""
    sh = torch.arange(32, device=bits.device, dtype=torch.int32)
    b = ((bits.unsqueeze(-1) >> sh) & 1).bool()
    return b.flatten(-2)[..., :hw]


def mask_pred(E, F_):
    E, F_ = _f32(E), _f32(F_)
    B, N, _ = E.shape
    H, W = F_.shape[2:]
    out = torch.empty((B, N, H, W), dtype=torch.float32, device=E.device)
    nat.check(nat.load().pn_mask_pred(E.data_ptr(), F_.data_ptr(), out.data_ptr(), B, N, H * W, _stream(E)),
              "pn_mask_pred")
    return out


def linear(x, weight, bias=None, relu=False, resid=None):
    x, weight = _f32(x), _f32(weight)
    M, K = x.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    nat.check(nat.load().pn_linear(x.data_ptr(), K, weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                   resid.data_ptr() if resid is not None else None, y.data_ptr(), N, M, N, K,
                                   int(relu), _stream(x)), "pn_linear")
    return y


def linear_tc(x, weight, bias=None, passes=3):
    """tcgen05 tensor-core linear (3xTF32 by default): x [M,K] @ weight[N,K]^T + bias."""
    x, weight = _f32(x), _f32(weight)
    M, K = x.shape
    N = weight.shape[0]
    lib = nat.load()
    need = lib.pn_linear_tc_workspace_bytes(M, N, K)
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    nat.check(lib.pn_linear_tc(x.data_ptr(), K, weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                               y.data_ptr(), N, M, N, K, passes, ws.data_ptr(), need, _stream(x)), "pn_linear_tc")
    return y


def add_layernorm(x, resid, gamma, beta):
    x = _f32(x)
    y = torch.empty_like(x)
    nat.check(nat.load().pn_add_layernorm(x.data_ptr(), resid.data_ptr() if resid is not None else None,
                                          gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), x.shape[0], _stream(x)),
              "pn_add_layernorm")
    return y


def mha_core(q, k, v, mask_bits=None, rowany=None):
    """q [B,Nq,256], k/v [B,Nk,256] (already projected) -> [B,Nq,256]."""
    q, k, v = _f32(q), _f32(k), _f32(v)
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    lib = nat.load()
    need = lib.pn_mha_workspace_bytes(B, Nq, Nk)
    ws = torch.empty(need, dtype=torch.uint8, device=q.device)
    out = torch.empty((B, Nq, 256), dtype=torch.float32, device=q.device)
    nat.check(lib.pn_mha_core(q.data_ptr(), 256, k.data_ptr(), 256, v.data_ptr(), 256,
                              mask_bits.data_ptr() if mask_bits is not None else None,
                              mask_bits.shape[-1] if mask_bits is not None else 0,
                              rowany.data_ptr() if rowany is not None else None, out.data_ptr(), B, Nq, Nk,
                              ws.data_ptr(), need, _stream(q)), "pn_mha_core")
    return out


def mha_core_tc(q, k, v, mask_bits=None, rowany=None):
    """tcgen05 flash attention (3xTF32): q [B,Nq,256], k/v [B,Nk,256] (already projected) -> [B,Nq,256]."""
    q, k, v = _f32(q), _f32(k), _f32(v)
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    lib = nat.load()
    need = lib.pn_mha_core_tc_workspace_bytes(B, Nq, Nk)
    ws = torch.empty(need, dtype=torch.uint8, device=q.device)
    out = torch.empty((B, Nq, 256), dtype=torch.float32, device=q.device)
    nat.check(lib.pn_mha_core_tc(q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                 mask_bits.data_ptr() if mask_bits is not None else None,
                                 mask_bits.shape[-1] if mask_bits is not None else 0,
                                 rowany.data_ptr() if rowany is not None else None, out.data_ptr(), B, Nq, Nk,
                                 ws.data_ptr(), need, _stream(q)), "pn_mha_core_tc")
    return out


def _conv_struct(conv):
    cv = nat.PnConvTiny()
    cv.mid_channels = conv.conv_layers[0][0].out_channels
    keep = []
    for i in range(3):
        c = conv.conv_layers[i][0]
        wt, b = _f32(c.weight.detach()), _f32(c.bias.detach())
        keep += [wt, b]
        cv.w[i], cv.b[i] = wt.data_ptr(), b.data_ptr()
    return cv, keep


def conv_tiny(x, conv):
    """x [B,N,N] -> ConvTiny(x) with ``conv`` any module exposing ``conv_layers.{0,1,2}.0``."""
    x = _f32(x)
    B, N, _ = x.shape
    cv, keep = _conv_struct(conv)
    lib = nat.load()
    need = lib.pn_ppn_workspace_bytes(B, N, 1, cv.mid_channels)
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    y = torch.empty_like(x)
    nat.check(lib.pn_conv_tiny(x.data_ptr(), C.byref(cv), y.data_ptr(), B, N, ws.data_ptr(), need, _stream(x)),
              "pn_conv_tiny")
    return y


def topk_pairs(importance, K, query=None):
    """importance [B,N,N] -> (idx, sub_pos, obj_pos [B,K] int64, pair_feat [B,2K,256] or None)."""
    imp = _f32(importance)
    B, N, _ = imp.shape
    i64 = dict(dtype=torch.int64, device=imp.device)
    idx, sp, op = torch.empty((B, K), **i64), torch.empty((B, K), **i64), torch.empty((B, K), **i64)
    pair = None
    if query is not None:
        query = _f32(query)
        pair = torch.empty((B, 2 * K, 256), dtype=torch.float32, device=imp.device)
    nat.check(nat.load().pn_topk_pairs(imp.data_ptr(), idx.data_ptr(), sp.data_ptr(), op.data_ptr(),
                                       query.data_ptr() if query is not None else None,
                                       pair.data_ptr() if pair is not None else None, B, N, K, _stream(imp)),
              "pn_topk_pairs")
    return idx, sp, op, pair


class PpnPlan:
    """Pre-allocated buffers for repeated ``pn_ppn_forward`` calls (micro-benchmark 5a/5b)."""

    def __init__(self, B, N, K, device, mid_channels=0):
        """mid_channels = 0: pair matrix + top-k only (5a); 64: workspace for ``run_embeds(..., conv=...)`` as well."""
        lib = nat.load()
        self.mid_channels = mid_channels
        self.B, self.N, self.K = B, N, K
        self.need = lib.pn_ppn_workspace_bytes(B, N, K, mid_channels)
        self.ws = torch.empty(self.need, dtype=torch.uint8, device=device)
        self.importance = torch.empty((B, N, N), dtype=torch.float32, device=device)
        self.sub_pos = torch.empty((B, K), dtype=torch.int64, device=device)
        self.obj_pos = torch.empty((B, K), dtype=torch.int64, device=device)
        self.idx = torch.empty((B, K), dtype=torch.int64, device=device)

    def run_embeds(self, sub_embed, obj_embed, conv=None, raw=None):
        """microbench mode: already-normalised embeddings -> pair matrix (-> conv) -> top-k."""
        cvp, keep = (None, None)
        if conv is not None:
            if self.mid_channels <= 0:
                raise ValueError("PpnPlan was sized without ConvTiny workspace (pass mid_channels=64)")
            cv, keep = _conv_struct(conv)
            cvp = C.byref(cv)
        nat.check(nat.load().pn_ppn_forward(sub_embed.data_ptr(), obj_embed.data_ptr(), None, None, cvp,
                                            raw.data_ptr() if raw is not None else None, self.importance.data_ptr(),
                                            self.idx.data_ptr(), self.sub_pos.data_ptr(), self.obj_pos.data_ptr(),
                                            None, self.B, self.N, self.K, self.ws.data_ptr(), self.need,
                                            _stream(sub_embed)), "pn_ppn_forward")
        return self.importance, self.idx, self.sub_pos, self.obj_pos

    def run_embeds_bf16(self, sub_embed, obj_embed):
        """bf16 microbench mode (``pn_ppn_pair_topk_bf16``): bf16 embeddings -> fp32 pair matrix + int64 top-k."""
        if sub_embed.dtype != torch.bfloat16 or obj_embed.dtype != torch.bfloat16:
            raise TypeError("run_embeds_bf16 takes torch.bfloat16 embeddings")
        if not (sub_embed.is_contiguous() and obj_embed.is_contiguous()):
            raise ValueError("embeddings must be contiguous [B,N,256]")
        nat.check(nat.load().pn_ppn_pair_topk_bf16(sub_embed.data_ptr(), obj_embed.data_ptr(),
                                                   self.importance.data_ptr(), self.idx.data_ptr(),
                                                   self.sub_pos.data_ptr(), self.obj_pos.data_ptr(), self.B, self.N,
                                                   self.K, self.ws.data_ptr(), self.need, _stream(sub_embed)),
                  "pn_ppn_pair_topk_bf16")
        return self.importance, self.idx, self.sub_pos, self.obj_pos


def gather_rows(src, idx):
    """src [B,Nsrc,...], idx [B,R] int64 -> [B,R,...]."""
    src = _f32(src)
    B, Nsrc = src.shape[:2]
    R = idx.shape[1]
    L = 1
    for s in src.shape[2:]:
        L *= s
    out = torch.empty((B, R) + tuple(src.shape[2:]), dtype=torch.float32, device=src.device)
    nat.check(nat.load().pn_gather_rows(src.data_ptr(), idx.contiguous().data_ptr(), out.data_ptr(), B, Nsrc, R, L,
                                        _stream(src)), "pn_gather_rows")
    return out
