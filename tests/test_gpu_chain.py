"""GPU (-m gpu): the Relation Fusion decoder through the C-ABI stage entry point ``pn_relation_fusion_forward`` in its
three variants (PN_OPT_FUSED_CHAIN):
  2 = the fused decoder-chain kernel (pairnet_b200/csrc/chain.cu): six BaseTransformerLayers + the relation classifier
      in ONE tcgen05 cluster launch (+ 4 key-side GEMM launches),
  1 = auto: chain from 8 images per call, below that per-op kernels with the key side and the cross attention on tcgen05,
  0 = round-1 per-op kernels (warp-MMA linears, FFMA attention),
against the CPU oracle's relation decoder evaluated in float64 on the same pair features (pairnet_head.py:353-378)."""
import ctypes as C

import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _build(N, R, seed=777):
    from oracle.head import HeadHyper, OCrossHead2
    from oracle.weights import fixture_state_dict
    from pairnet_b200.registry import build_head
    from tests.util import product_head_cfg
    o = OCrossHead2(HeadHyper(num_obj_query=N, num_rel_query=R, with_pixel_decoder=False))
    o.load_state_dict(fixture_state_dict(o, seed))
    o.eval()
    cfg = product_head_cfg()
    cfg.update(pixel_decoder=None, num_obj_query=N, num_rel_query=R)
    p = build_head(cfg)
    p.load_state_dict(o.state_dict(), strict=True)
    return o, p.cuda().eval()


def _oracle_relation_fusion(o, pair_feat):
    """pair_feat [B,2K,256] -> (rel_feat [B,R,256], rel_preds [B,R,56]) in float64 (oracle/head.py relation loop)."""
    import copy
    o = copy.deepcopy(o).double()
    B = pair_feat.shape[0]
    pf = pair_feat.double().transpose(0, 1)
    x = o.rel_query_feat.weight.unsqueeze(1).repeat((1, B, 1))
    qe = o.rel_query_embed.weight.unsqueeze(1).repeat((1, B, 1))
    ke = o.rel_query_embed2.weight.unsqueeze(1).repeat((1, B, 1))
    ve = o.rel_query_embed3.weight.unsqueeze(1).repeat((1, B, 1))
    with torch.no_grad():
        for layer in o.relation_decoder.layers:
            x = layer(query=x, key=pf, value=pf, query_pos=qe, key_pos=ke, value_pos=ve)
        return x.transpose(0, 1), o.rel_cls_embed(x.transpose(0, 1))


def _run_stage(p, pair_feat, mode):
    from pairnet_b200 import _native as nat
    lib = nat.load()
    w = p.native_weights()
    B, K2, _ = pair_feat.shape
    R = p.num_rel_query
    need = lib.pn_relation_fusion_workspace_bytes(B, R, K2, 2048)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    rel = torch.zeros((B, R, p.num_relations), device="cuda")
    feat = torch.zeros((B, R, 256), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib.pn_set_option(nat.PN_OPT_FUSED_CHAIN, int(mode))
    try:
        nat.check(lib.pn_relation_fusion_forward(C.byref(w.rel), pair_feat.data_ptr(), rel.data_ptr(), feat.data_ptr(), B,
                                                 K2, ws.data_ptr(), need, st), "pn_relation_fusion_forward")
        torch.cuda.synchronize()
        launches = lib.pn_last_launch_count()
    finally:
        lib.pn_set_option(nat.PN_OPT_FUSED_CHAIN, 1)
    return rel, feat, launches


@pytest.mark.parametrize("N,R,B", [(100, 100, 1), (100, 100, 2), (50, 30, 3), (200, 200, 1), (100, 100, 5), (100, 100, 9)])
def test_relation_fusion_variants(N, R, B):
    o, p = _build(N, R)
    g = torch.Generator().manual_seed(1234 + R + B)
    pair = torch.randn((B, 2 * R, 256), generator=g)
    ref_feat, ref_rel = _oracle_relation_fusion(o, pair)
    rel_c, feat_c, n_c = _run_stage(p, pair.cuda(), 2)
    rel_a, feat_a, n_a = _run_stage(p, pair.cuda(), 1)
    rel_u, feat_u, n_u = _run_stage(p, pair.cuda(), 0)
    # the whole Relation Fusion decoder in <= 12 launches on the chain (per-op paths: ~85-95)
    assert n_c <= 12, n_c
    assert n_u > 50, n_u
    assert (n_a <= 12) == (B >= 8), (n_a, B)  # auto mode takes the chain from 8 images
    for name, rel, feat in (("chain", rel_c, feat_c), ("auto", rel_a, feat_a), ("per-op", rel_u, feat_u)):
        assert torch.isfinite(rel).all() and torch.isfinite(feat).all(), name
        assert rel_err(feat, ref_feat) < 2e-5, f"{name}: rel_feat vs float64 oracle"
        assert rel_err(rel, ref_rel) < 2e-5, f"{name}: rel logits vs float64 oracle"


def test_relation_fusion_is_deterministic():
    o, p = _build(100, 100)
    pair = torch.randn((2, 200, 256), generator=torch.Generator().manual_seed(5)).cuda()
    for mode in (2, 1):
        a, fa, _ = _run_stage(p, pair, mode)
        for _ in range(3):
            b, fb, _ = _run_stage(p, pair, mode)
            assert torch.equal(a, b) and torch.equal(fa, fb)
