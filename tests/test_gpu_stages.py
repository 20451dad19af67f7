"""GPU (-m gpu): stage-wise parity of every C-ABI entry point against the CPU oracle / torch fp64 on the
same seeded inputs.  Bit-exact for index/byte work, stated tolerance for fp32."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _t(shape, seed, scale=1.0):
    from oracle.weights import numpy_tensor
    return numpy_tensor(shape, seed, scale)


@pytest.mark.parametrize("h,w", [(4, 6), (13, 17), (25, 42), (50, 84)])
def test_sine_posenc(h, w):
    from oracle.bricks import sine_positional_encoding
    from pairnet_b200 import ops
    ref = sine_positional_encoding(torch.zeros(1, h, w, dtype=torch.bool))[0].flatten(1).t()
    got = ops.sine_posenc(h, w).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 5e-6  # sinf/cosf/powf ulp differences only


def test_level_prep_bit_exact():
    from pairnet_b200 import ops
    B, h, w = 2, 9, 13
    mem, lvl, pos = _t((B, 256, h, w), 1), _t((256,), 2), _t((h * w, 256), 3)
    x, xp = ops.level_prep(mem.cuda(), lvl.cuda(), pos.cuda())
    rx = mem.flatten(2).permute(0, 2, 1) + lvl.view(1, 1, -1)
    assert torch.equal(x.cpu(), rx)
    assert torch.equal(xp.cpu(), rx + pos[None])


@pytest.mark.parametrize("H,W,h,w", [(32, 48, 4, 6), (40, 56, 20, 28), (50, 84, 13, 21), (200, 334, 25, 42)])
def test_mask_feature_resize(H, W, h, w):
    from pairnet_b200 import ops
    Fm = _t((1, 256, H, W), 4)
    ref = F.interpolate(Fm, (h, w), mode="bilinear", align_corners=False).flatten(2)
    got = ops.mask_feature_resize(Fm.cuda(), h, w).cpu()
    assert got.shape[2] % 64 == 0
    assert float((got[:, :, : h * w] - ref).abs().max()) < 2e-6 * float(ref.abs().max()) + 1e-6
    assert float(got[:, :, h * w:].abs().max()) == 0.0 if got.shape[2] > h * w else True


@pytest.mark.parametrize("M,N,K,relu,resid", [(200, 256, 256, False, False), (200, 2048, 256, True, False),
                                             (200, 256, 2048, False, True), (200, 134, 256, False, False),
                                             (37, 56, 256, False, False), (2100, 256, 256, False, False),
                                             (4099, 512, 256, True, True)])
def test_linear(M, N, K, relu, resid):
    from pairnet_b200 import ops
    x, w, b = _t((M, K), 5), _t((N, K), 6, 0.1), _t((N,), 7)
    r = _t((M, N), 8) if resid else None
    ref = x.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.relu()
    if resid:
        ref = ref + r.double()
    got = ops.linear(x.cuda(), w.cuda(), b.cuda(), relu=relu, resid=r.cuda() if resid else None).cpu()
    assert rel_err(got, ref) < 2e-6


def test_add_layernorm():
    from pairnet_b200 import ops
    x, r, g, b = _t((203, 256), 9), _t((203, 256), 10), _t((256,), 11), _t((256,), 12)
    ref = F.layer_norm((x + r).double(), (256,), g.double(), b.double(), 1e-5)
    got = ops.add_layernorm(x.cuda(), r.cuda(), g.cuda(), b.cuda()).cpu()
    assert rel_err(got, ref) < 2e-6


def _mha_ref(q, k, v, blocked=None):
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    qh = q.double().view(B, Nq, 8, 32).transpose(1, 2) / np.sqrt(32.0)
    kh = k.double().view(B, Nk, 8, 32).transpose(1, 2)
    vh = v.double().view(B, Nk, 8, 32).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if blocked is not None:
        blocked = blocked.clone()
        blocked[blocked.all(-1)] = False  # pairnet_head.py:300
        s = s.masked_fill(blocked[:, None], float("-inf"))
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Nq, 256)


@pytest.mark.parametrize("Nq,Nk", [(100, 100), (100, 200), (100, 1050), (100, 4200), (130, 333)])
def test_mha_core_unmasked(Nq, Nk):
    from pairnet_b200 import ops
    q, k, v = _t((2, Nq, 256), 13), _t((2, Nk, 256), 14), _t((2, Nk, 256), 15)
    got = ops.mha_core(q.cuda(), k.cuda(), v.cuda()).cpu()
    assert rel_err(got, _mha_ref(q, k, v)) < 5e-6


@pytest.mark.parametrize("Nk", [70, 1050, 4200])
def test_mha_core_masked_with_fully_blocked_rows(Nk):
    from pairnet_b200 import ops
    B, Nq = 2, 100
    q, k, v = _t((B, Nq, 256), 16, 2.0), _t((B, Nk, 256), 17), _t((B, Nk, 256), 18)
    rng = np.random.default_rng(Nk)
    blocked = torch.from_numpy(rng.random((B, Nq, Nk)) < 0.6)
    blocked[0, 3] = True            # fully blocked row -> must attend everywhere
    blocked[1, 99] = True
    blocked[0, 5] = True
    blocked[0, 5, Nk - 1] = False   # a single open key at the very end
    blocked[1, 7, 64:] = True       # open only inside the first tile
    words = (Nk + 63) // 64 * 2
    padded = torch.ones((B, Nq, words * 32), dtype=torch.bool)
    padded[:, :, :Nk] = blocked
    bits = (padded.view(B, Nq, words, 32).long() << torch.arange(32)).sum(-1)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
    rowany = (~blocked).any(-1).to(torch.int32).flatten()
    got = ops.mha_core(q.cuda(), k.cuda(), v.cuda(), bits.cuda(), rowany.cuda()).cpu()
    assert rel_err(got, _mha_ref(q, k, v, blocked)) < 5e-6


@pytest.mark.parametrize("Nq,Nk", [(100, 128), (100, 1050), (100, 4200), (130, 333), (200, 2100)])
def test_mha_core_tc_unmasked(Nq, Nk):
    """tcgen05 flash attention (QK^T and PV on tensor cores, P through TMEM) vs fp64 math."""
    from pairnet_b200 import ops
    q, k, v = _t((2, Nq, 256), 13), _t((2, Nk, 256), 14), _t((2, Nk, 256), 15)
    got = ops.mha_core_tc(q.cuda(), k.cuda(), v.cuda()).cpu()
    assert rel_err(got, _mha_ref(q, k, v)) < 1e-5


@pytest.mark.parametrize("Nk", [200, 1050, 4200])
def test_mha_core_tc_masked_with_fully_blocked_rows(Nk):
    from pairnet_b200 import ops
    B, Nq = 2, 100
    q, k, v = _t((B, Nq, 256), 16, 2.0), _t((B, Nk, 256), 17), _t((B, Nk, 256), 18)
    rng = np.random.default_rng(Nk)
    blocked = torch.from_numpy(rng.random((B, Nq, Nk)) < 0.6)
    blocked[0, 3] = True
    blocked[1, 99] = True
    blocked[0, 5] = True
    blocked[0, 5, Nk - 1] = False
    blocked[1, 7, 64:] = True
    blocked[1, 8, :Nk - 130] = True  # open keys only in the last two tiles
    words = (Nk + 63) // 64 * 2
    padded = torch.ones((B, Nq, words * 32), dtype=torch.bool)
    padded[:, :, :Nk] = blocked
    bits = (padded.view(B, Nq, words, 32).long() << torch.arange(32)).sum(-1)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
    rowany = (~blocked).any(-1).to(torch.int32).flatten()
    got = ops.mha_core_tc(q.cuda(), k.cuda(), v.cuda(), bits.cuda(), rowany.cuda()).cpu()
    assert rel_err(got, _mha_ref(q, k, v, blocked)) < 1e-5


@pytest.mark.parametrize("hw", [24, 1050, 4200])
def test_attn_mask_bits(hw):
    from pairnet_b200 import ops
    B, N = 2, 100
    E = _t((B, N, 256), 19)
    ldf = (hw + 63) // 64 * 64
    Fl = torch.zeros((B, 256, ldf))
    Fl[:, :, :hw] = _t((B, 256, hw), 20)
    E[0, 4] = 0.0  # all-zero row: logits == 0 -> not < 0 -> every key open
    Fl[1, :, :hw] = torch.where(Fl[1, :, :hw] > 0, Fl[1, :, :hw], -Fl[1, :, :hw])
    E[1, 9] = -E[1, 9].abs()  # negative embed x positive features -> all blocked -> rowany 0
    bits, rowany = ops.attn_mask_bits(E.cuda(), Fl.cuda(), hw)
    got = ops.unpack_bits(bits, hw).cpu()
    logits = torch.einsum("bqc,bcp->bqp", E.double(), Fl[:, :, :hw].double())
    ref = logits < 0
    sure = logits.abs() > 1e-4
    assert torch.equal(got[sure], ref[sure])
    assert (got != ref).float().mean() < 1e-5
    assert not got[0, 4].any() and got[1, 9].all()
    ra = rowany.cpu().view(B, N)
    assert ra[1, 9] == 0 and ra[0, 4] == 1
    full = ops.unpack_bits(bits, ldf).cpu()
    assert full[:, :, hw:].all()  # padding keys are blocked


@pytest.mark.parametrize("hw,N,B", [(24, 100, 2), (1050, 100, 2), (4200, 100, 2), (700, 200, 1), (333, 37, 5)])
def test_attn_mask_bits_tc(hw, N, B):
    """tensor-core mask bits (tcgen05 3xTF32 GEMM, keys as M tiles, sign + warp-ballot epilogue) on token-major
    features: same contract as the FFMA kernel."""
    from pairnet_b200 import _native as nat, ops
    lib = nat.load()
    E = _t((B, N, 256), 19)
    Ft = _t((B, hw, 256), 20)                       # token-major resized mask features
    E[0, 4] = 0.0
    Ft[B - 1] = Ft[B - 1].abs()
    E[B - 1, 9] = -E[B - 1, 9].abs()
    words = (hw + 63) // 64 * 2
    Ec, Fc = E.cuda(), Ft.cuda()
    bits = torch.zeros((B, N, words), dtype=torch.int32, device="cuda")
    rowany = torch.zeros((B * N,), dtype=torch.int32, device="cuda")
    need = lib.pn_mask_tc_workspace_bytes(B, N)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    nat.check(lib.pn_attn_mask_bits_tc(Ec.data_ptr(), Fc.data_ptr(), bits.data_ptr(), rowany.data_ptr(), B, N, hw, words,
                                       ws.data_ptr(), need, torch.cuda.current_stream().cuda_stream), "bits_tc")
    got = ops.unpack_bits(bits, hw).cpu()
    logits = torch.einsum("bqc,bpc->bqp", E.double(), Ft.double())
    ref = logits < 0
    sure = logits.abs() > 1e-4
    assert torch.equal(got[sure], ref[sure])
    assert (got != ref).float().mean() < 1e-5
    assert not got[0, 4].any() and got[B - 1, 9].all()
    ra = rowany.cpu().view(B, N)
    assert ra[B - 1, 9] == 0 and ra[0, 4] == 1
    assert torch.equal(ra.bool(), ~got.all(-1))
    full = ops.unpack_bits(bits, words * 32).cpu()
    assert full[:, :, hw:].all()  # padding keys are blocked


@pytest.mark.parametrize("B,N,H,W", [(2, 100, 23, 31), (1, 200, 40, 56), (3, 37, 16, 24)])
def test_mask_pred_tc_and_token_layout_helpers(B, N, H, W):
    """final mask_pred on the tcgen05 GEMM with a transposed store; NCHW -> token-major copy is exact."""
    from pairnet_b200 import _native as nat
    lib = nat.load()
    st = torch.cuda.current_stream().cuda_stream
    E, Fm = _t((B, N, 256), 21), _t((B, 256, H, W), 22)
    Ec, Fc = E.cuda(), Fm.cuda()
    Ft = torch.empty((B, H * W, 256), device="cuda")
    nat.check(lib.pn_nchw_to_tokens(Fc.data_ptr(), Ft.data_ptr(), B, H * W, st), "nchw_to_tokens")
    assert torch.equal(Ft.cpu(), Fm.flatten(2).transpose(1, 2))
    need = lib.pn_mask_tc_workspace_bytes(B, N)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    out = torch.full((B, N, H, W), float("nan"), device="cuda")
    nat.check(lib.pn_mask_pred_tc(Ec.data_ptr(), Ft.data_ptr(), out.data_ptr(), B, N, H * W, ws.data_ptr(), need, st),
              "mask_pred_tc")
    # 3xTF32 (hi/lo split operands on the tensor pipe): fp32-level, slightly above the FFMA kernel's 2e-6
    assert rel_err(out.cpu(), torch.einsum("bqc,bchw->bqhw", E.double(), Fm.double())) < 5e-6


@pytest.mark.parametrize("H,W,h,w", [(32, 48, 4, 6), (40, 56, 20, 28), (200, 334, 25, 42)])
def test_mask_feature_resize_tokens(H, W, h, w):
    from pairnet_b200 import _native as nat
    lib = nat.load()
    Fm = _t((2, 256, H, W), 4)
    ref = F.interpolate(Fm, (h, w), mode="bilinear", align_corners=False).flatten(2).transpose(1, 2)
    Ft = Fm.cuda().contiguous(memory_format=torch.channels_last)
    out = torch.empty((2, h * w, 256), device="cuda")
    nat.check(lib.pn_mask_feature_resize_tokens(Ft.data_ptr(), out.data_ptr(), 2, H, W, h, w,
                                                torch.cuda.current_stream().cuda_stream), "resize_tokens")
    assert float((out.cpu() - ref).abs().max()) < 2e-6 * float(ref.abs().max()) + 1e-6


def test_mask_pred():
    from pairnet_b200 import ops
    E, Fm = _t((2, 100, 256), 21), _t((2, 256, 23, 31), 22)
    got = ops.mask_pred(E.cuda(), Fm.cuda()).cpu()
    assert rel_err(got, torch.einsum("bqc,bchw->bqhw", E.double(), Fm.double())) < 2e-6


def test_conv_tiny_against_reference_golden():
    """weights/inputs regenerated from seeds; expected output produced by the REFERENCE ConvTiny."""
    from oracle.head import OConvTiny
    from oracle.make_golden import CONV_CASES
    from oracle.weights import numpy_state_dict, numpy_tensor
    from pairnet_b200 import ops
    for tag, mid, B, N, seed in CONV_CASES:
        g = np.load(os.path.join(GOLDEN, f"convtiny_ref_{tag}.npz"))
        m = OConvTiny(mid_channels=mid)
        m.load_state_dict(numpy_state_dict(m, seed))
        x = torch.tanh(numpy_tensor((B, N, N), seed + 100))
        got = ops.conv_tiny(x.cuda(), m.cuda()).cpu()
        assert rel_err(got, g["out"]) < 1e-5, tag


@pytest.mark.parametrize("B,N", [(2, 100), (3, 37), (1, 130), (2, 200), (5, 16), (1, 8)])
def test_conv_tiny_tcgen05_matches_fp64_and_ffma(B, N):
    """conv2 as a tcgen05 implicit GEMM (shifted-window descriptors over zero-padded TMA slabs, 3xTF32) against the
    fp64 convolution of the same module and against the exact-fp32 FFMA kernels; ragged patches and image edges."""
    from oracle.head import OConvTiny
    from pairnet_b200 import _native as nat, ops
    torch.manual_seed(40 + N)
    m = OConvTiny(mid_channels=64)
    x = torch.tanh(_t((B, N, N), 300 + N))
    with torch.no_grad():
        ref = m.double()(x.double())
    m = m.float().cuda()
    lib = nat.load()
    assert lib.pn_get_option(nat.PN_OPT_CONV_TC) == 1
    got = ops.conv_tiny(x.cuda(), m).cpu()
    lib.pn_set_option(nat.PN_OPT_CONV_TC, 0)
    try:
        ffma = ops.conv_tiny(x.cuda(), m).cpu()
    finally:
        lib.pn_set_option(nat.PN_OPT_CONV_TC, 1)
    scale = float(ref.abs().max())
    print(f"conv_tiny B={B} N={N}: tcgen05 {float((got.double() - ref).abs().max()) / scale:.2e}  "
          f"ffma {float((ffma.double() - ref).abs().max()) / scale:.2e} (max abs err / max |ref|, vs fp64)")
    assert float((ffma.double() - ref).abs().max()) < 1e-5 * scale
    assert float((got.double() - ref).abs().max()) < 1e-5 * scale
    assert float((got - ffma).abs().max()) < 1e-5 * scale


@pytest.mark.parametrize("N,K", [(100, 100), (200, 100), (37, 5), (100, 1), (64, 1024), (400, 100)])
def test_topk_pairs_bit_exact(N, K):
    from pairnet_b200 import ops
    B = 3
    imp = _t((B, N, N), 23 + N)
    q = _t((B, N, 256), 24)
    idx, sp, op, pair = ops.topk_pairs(imp.cuda(), K, q.cuda())
    ref_v, ref_i = torch.topk(imp.flatten(1), K)
    assert torch.equal(idx.cpu(), ref_i)  # tie-free continuous values: exact, in order
    assert torch.equal(sp.cpu(), torch.div(ref_i, N, rounding_mode="trunc"))
    assert torch.equal(op.cpu(), torch.remainder(ref_i, N))
    exp = torch.cat([torch.gather(q, 1, sp.cpu()[..., None].expand(-1, -1, 256)),
                     torch.gather(q, 1, op.cpu()[..., None].expand(-1, -1, 256))], 1)
    assert torch.equal(pair.cpu(), exp)


def test_topk_pairs_ties_negative_and_special_values():
    from oracle.head import stable_topk
    from pairnet_b200 import ops
    N, K = 50, 100
    rng = np.random.default_rng(3)
    imp = torch.from_numpy(rng.integers(-3, 4, size=(2, N, N)).astype(np.float32))  # massive ties
    imp[0, 0, 0] = float("inf")
    imp[0, 7, 7] = -0.0
    imp[1] = -imp[1].abs() - 1.0  # all negative
    idx, sp, op, _ = ops.topk_pairs(imp.cuda(), K)
    for b in range(2):
        assert idx[b].cpu().tolist() == stable_topk(imp[b].flatten().numpy(), K).tolist()
    const = torch.full((1, N, N), 0.25)
    idx, _, _, _ = ops.topk_pairs(const.cuda(), K)
    assert idx[0].cpu().tolist() == list(range(K))  # all equal -> lowest flat indices, ascending


def test_gather_rows_bit_exact():
    from pairnet_b200 import ops
    src = _t((2, 100, 134), 25)
    idx = torch.from_numpy(np.random.default_rng(1).integers(0, 100, size=(2, 100)))
    assert torch.equal(ops.gather_rows(src.cuda(), idx.cuda()).cpu(),
                       torch.gather(src, 1, idx[..., None].expand(-1, -1, 134)))
    seg = _t((2, 100, 20, 28), 26)
    assert torch.equal(ops.gather_rows(seg.cuda(), idx.cuda()).cpu(),
                       torch.gather(seg, 1, idx[..., None, None].expand(-1, -1, 20, 28)))


def test_error_conventions():
    from pairnet_b200 import _native as nat, ops
    with pytest.raises(nat.NativeError):  # K > N*N
        ops.topk_pairs(torch.zeros(1, 4, 4, device="cuda"), 100)
    with pytest.raises(nat.NativeError):  # K not a multiple of the GEMM k-tile
        ops.linear(torch.zeros(8, 100, device="cuda"), torch.zeros(8, 100, device="cuda"))
    assert b"gemm" in nat.load().pn_last_error_string()


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 256, 256), (1000, 256, 256), (4200, 512, 256), (333, 384, 1024)])
def test_linear_tc_3xtf32_matches_fp32(M, N, K):
    """tcgen05 GEMM (TMA + UMMA kind::tf32, hi/lo split): fp32-level accuracy; single pass = TF32 accuracy."""
    from pairnet_b200 import ops
    x, w, b = _t((M, K), 30), _t((N, K), 31, 0.1), _t((N,), 32)
    ref = x.double() @ w.double().t() + b.double()
    got3 = ops.linear_tc(x.cuda(), w.cuda(), b.cuda(), passes=3).cpu()
    assert rel_err(got3, ref) < 2e-5  # tensor-core fp32 accumulation over K up to 1024
    got1 = ops.linear_tc(x.cuda(), w.cuda(), b.cuda(), passes=1).cpu()
    e1 = rel_err(got1, ref)
    assert 1e-5 < e1 < 5e-3  # really went through TF32 tensor cores


@pytest.mark.parametrize("M,N,K", [(1500, 256, 256), (4097, 288, 256), (2100, 1024, 256), (1111, 256, 1024), (300, 64, 64)])
def test_linear_tc_bf16x3(M, N, K):
    """`pn_linear_tc_bf16x3`: fp32 in / out GEMM as three bf16 products on tcgen05 kind::f16 (hi*hi + hi*lo + lo*hi).
    Error model: dropped lo*lo and third-digit terms, 2^-16..2^-17 per product -> bound 5e-5 of the output scale; the
    bf16 split itself is checked exactly (hi = bf16(w), lo = bf16(w - hi))."""
    from pairnet_b200 import _native as nat
    lib = nat.load()
    st = torch.cuda.current_stream().cuda_stream
    x, w, b = _t((M, K), 30).cuda(), _t((N, K), 31, 0.1).cuda(), _t((N,), 32).cuda()
    wh = torch.empty((N, K), dtype=torch.bfloat16, device="cuda")
    wl = torch.empty((N, K), dtype=torch.bfloat16, device="cuda")
    nat.check(lib.pn_split_bf16(w.data_ptr(), wh.data_ptr(), wl.data_ptr(), w.numel(), st), "pn_split_bf16")
    assert torch.equal(wh, w.to(torch.bfloat16)) and torch.equal(wl, (w - wh.float()).to(torch.bfloat16))
    y = torch.empty((M, N), device="cuda")
    nat.check(lib.pn_linear_tc_bf16x3(x.data_ptr(), wh.data_ptr(), wl.data_ptr(), b.data_ptr(), y.data_ptr(), N, M, N, K,
                                      st), "pn_linear_tc_bf16x3")
    ref = x.double() @ w.double().t() + b.double()
    err = rel_err(y, ref)
    assert 1e-7 < err < 5e-5, err
    # the model of the arithmetic: exact products of the bf16 digits, fp32-accumulated
    xh = x.to(torch.bfloat16).double(); xl = (x - xh.float()).to(torch.bfloat16).double()
    model = xh @ wh.double().t() + xh @ wl.double().t() + xl @ wh.double().t() + b.double()
    assert rel_err(y, model) < 5e-6


def test_linear_tc_bf16x3_rejects_what_it_cannot_run():
    """No silent fallback: K not a multiple of the 64-channel k-block, an unaligned bias, or the TMA-store epilogue switched
    off are errors with a message."""
    from pairnet_b200 import _native as nat
    lib = nat.load()
    st = torch.cuda.current_stream().cuda_stream
    x = torch.zeros(256, 96, device="cuda"); wh = torch.zeros(128, 96, dtype=torch.bfloat16, device="cuda")
    y = torch.zeros(256, 128, device="cuda"); b = torch.zeros(132, device="cuda")
    assert lib.pn_linear_tc_bf16x3(x.data_ptr(), wh.data_ptr(), wh.data_ptr(), None, y.data_ptr(), 128, 256, 128, 96, st) != 0
    assert b"K" in lib.pn_last_error_string()
    x = torch.zeros(256, 128, device="cuda"); wh = torch.zeros(128, 128, dtype=torch.bfloat16, device="cuda")
    assert lib.pn_linear_tc_bf16x3(x.data_ptr(), wh.data_ptr(), wh.data_ptr(), b.data_ptr() + 4, y.data_ptr(), 128, 256, 128,
                                   128, st) != 0                                   # bias not 16-byte aligned
    lib.pn_set_option(nat.PN_OPT_UMMA_TMA_STORE, 0)
    try:
        assert lib.pn_linear_tc_bf16x3(x.data_ptr(), wh.data_ptr(), wh.data_ptr(), None, y.data_ptr(), 128, 256, 128, 128,
                                       st) != 0
    finally:
        lib.pn_set_option(nat.PN_OPT_UMMA_TMA_STORE, 1)
    assert lib.pn_linear_tc_bf16x3(x.data_ptr(), wh.data_ptr(), wh.data_ptr(), None, y.data_ptr(), 128, 256, 128, 128, st) == 0
    torch.cuda.synchronize()
    assert float(y.abs().max()) == 0.0


def _pixel_decoder(seed=3):
    import os
    from pairnet_b200.registry import Config, PLUGIN_LAYERS
    from tests.util import ROOT
    cfg = Config.fromfile(os.path.join(ROOT, "configs", "pairnet_r50_b200.py"), import_custom_modules=False)
    pd = dict(cfg.model.bbox_head.pixel_decoder)
    pd.update(in_channels=[256, 512, 1024, 2048], feat_channels=256, out_channels=256)
    torch.manual_seed(seed)
    m = PLUGIN_LAYERS.build(pd)
    m.init_weights()
    with torch.no_grad():  # non-trivial offsets / attention logits (mmcv init zeroes those weights)
        for layer in m.encoder.layers:
            a = layer.attentions[0]
            a.sampling_offsets.weight.normal_(0, 0.02)
            a.attention_weights.weight.normal_(0, 0.05)
    return m.cuda().eval()


def test_msda_encoder_native_vs_torch():
    """§8f-1 upstream row: native deformable encoder (tcgen05 GEMMs + sampling kernel) vs the PyTorch
    grid_sample restatement on the same device, and vs the CPU oracle's pixel decoder."""
    from oracle.bricks import OMSDeformAttnPixelDecoder
    m = _pixel_decoder()
    feats = [_t((2, 256, 40, 56), 41), _t((2, 512, 20, 28), 42), _t((2, 1024, 10, 14), 43), _t((2, 2048, 5, 7), 44)]
    with torch.no_grad():
        m.encoder_impl = "native"
        mf_n, mem_n = m([f.cuda() for f in feats])
        m.encoder_impl = "torch"
        mf_t, mem_t = m([f.cuda() for f in feats])
        o = OMSDeformAttnPixelDecoder().eval()
        o.load_state_dict(m.state_dict())
        mf_o, mem_o = o(feats)
    for a, b in zip(mem_n, mem_t):
        assert rel_err(a, b) < 2e-4
    for a, b in zip(mem_n, mem_o):
        assert rel_err(a, b) < 2e-3  # cuDNN input convs run TF32 on the GPU (upstream plumbing), CPU oracle is fp32
    # mask_feature passes through cuDNN convs that run TF32 by default on the GPU (upstream plumbing)
    assert rel_err(mf_n, mf_t) < 1e-3
    assert rel_err(mf_n, mf_o) < 5e-3


def test_msda_encoder_bf16x3_vs_3xtf32():
    """PN_OPT_ENC_BF16X3 (default on): the encoder's GEMMs as three bf16 products on tcgen05 kind::f16 (raw fp32
    activations split in the SM, prepared bf16 weight planes) vs the 3xTF32 variant of the same kernel: ~1e-5 of the
    scale (stated bound 1e-4, two orders below the TF32 convolutions that feed the encoder), at a size that covers
    ragged M tiles, N = 288 (three n-tiles, the last one partial) and K = 1024."""
    from pairnet_b200 import _native as nat
    lib = nat.load()
    m = _pixel_decoder()
    feats = [_t((2, 256, 72, 100), 41).cuda(), _t((2, 512, 36, 50), 42).cuda(), _t((2, 1024, 18, 25), 43).cuda(),
             _t((2, 2048, 9, 13), 44).cuda()]
    assert lib.pn_get_option(nat.PN_OPT_ENC_BF16X3) == 1
    with torch.no_grad():
        mf_b, mem_b = m(feats)
        mf_b, mem_b = mf_b.clone(), [t.clone() for t in mem_b]
        lib.pn_set_option(nat.PN_OPT_ENC_BF16X3, 0)
        try:
            mf_t, mem_t = m(feats)
        finally:
            lib.pn_set_option(nat.PN_OPT_ENC_BF16X3, 1)
        m.encoder_impl = "torch"
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            _, mem_ref = m(feats)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
    for a, b, r in zip(mem_b, mem_t, mem_ref):
        assert a.shape == b.shape and not torch.equal(a, b)      # the two variants are different arithmetic
        assert rel_err(a, b) < 1e-4
        assert rel_err(a, r) < 3e-4 and rel_err(b, r) < 3e-4      # both agree with the fp32 PyTorch encoder
    assert rel_err(mf_b, mf_t) < 1e-3                             # downstream of a TF32 cuDNN 3x3 convolution


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("relu", [False, True])
def test_group_norm_native(channels_last, relu):
    import ctypes as C
    from pairnet_b200 import _native as nat
    lib = nat.load()
    B, H, W = 2, 37, 53
    x = _t((B, 256, H, W), 50, 3.0) + 1.5
    g, b = _t((256,), 51), _t((256,), 52)
    ref = F.group_norm(x.double(), 32, g.double(), b.double(), 1e-5)
    if relu:
        ref = ref.relu()
    xc = x.cuda()
    if channels_last:
        xc = xc.contiguous(memory_format=torch.channels_last)
    need = lib.pn_group_norm_workspace_bytes(B, H * W, 32)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    y = torch.empty_like(xc)
    gc, bc = g.cuda(), b.cuda()  # keep the device tensors alive while their pointers are in use
    nat.check(lib.pn_group_norm(xc.data_ptr(), gc.data_ptr(), bc.data_ptr(), y.data_ptr(), B, H * W, 32,
                                int(relu), int(channels_last), 1e-5, ws.data_ptr(), need,
                                torch.cuda.current_stream().cuda_stream), "gn")
    assert rel_err(y.cpu(), ref) < 5e-6


def test_level_prep_tokens_bit_exact_with_nchw_path():
    """token-major memories (level slices of the encoder output, batch stride > hw*256) give the same X / XP bits."""
    from pairnet_b200 import _native as nat
    lib = nat.load()
    B, h, w, nq = 2, 9, 13, 9 * 13 + 40
    enc = _t((B, nq, 256), 61).cuda()                      # [B,nq,256]; the level starts at token 24
    tok = enc[:, 24:24 + h * w]                           # strided view, batch stride nq*256
    lvl, pos = _t((256,), 62).cuda(), _t((h * w, 256), 63).cuda()
    x, xp = torch.empty((B, h * w, 256), device="cuda"), torch.empty((B, h * w, 256), device="cuda")
    nat.check(lib.pn_level_prep_tokens(tok.data_ptr(), tok.stride(0), lvl.data_ptr(), pos.data_ptr(), x.data_ptr(),
                                       xp.data_ptr(), B, h * w, torch.cuda.current_stream().cuda_stream), "lp")
    rx = tok + lvl.view(1, 1, -1)
    assert torch.equal(x, rx)
    assert torch.equal(xp, rx + pos[None])


@pytest.mark.parametrize("H,W,h,w", [(40, 56, 20, 28), (50, 84, 25, 42), (37, 53, 19, 27)])
def test_gn_upsample_add(H, W, h, w):
    """FPN merge: GN(lateral) + bilinear_up(top), top token-major with a batch stride (mmdet pixel decoder forward)."""
    from pairnet_b200 import _native as nat
    lib = nat.load()
    B = 2
    x = _t((B, 256, H, W), 64, 3.0) + 0.5
    g, b = _t((256,), 65), _t((256,), 66)
    enc = _t((B, h * w + 11, 256), 67)
    top = enc[:, 11:].transpose(1, 2).reshape(B, 256, h, w)
    ref = F.group_norm(x.double(), 32, g.double(), b.double(), 1e-5) + F.interpolate(
        top.double(), size=(H, W), mode="bilinear", align_corners=False)
    ref32 = F.group_norm(x, 32, g, b, 1e-5) + F.interpolate(top, size=(H, W), mode="bilinear", align_corners=False)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    encc, gc, bc = enc.cuda(), g.cuda(), b.cuda()
    topc = encc[:, 11:]
    need = lib.pn_group_norm_workspace_bytes(B, H * W, 32)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    nat.check(lib.pn_gn_upsample_add(xc.data_ptr(), gc.data_ptr(), bc.data_ptr(), topc.data_ptr(), topc.stride(0),
                                     xc.data_ptr(), B, H, W, h, w, 32, 1e-5, ws.data_ptr(), need,
                                     torch.cuda.current_stream().cuda_stream), "gn_upadd")
    assert rel_err(xc.cpu(), ref) < 5e-6
    assert rel_err(xc.cpu(), ref32) < 5e-6


@pytest.mark.parametrize("B,H,W", [(2, 40, 56), (1, 13, 20), (3, 25, 44)])
def test_conv1x1_nhwc_to_nchw(B, H, W):
    """mask_feature 1x1 conv as a tcgen05 3xTF32 GEMM with a transposed (NCHW) store: fp32-level accuracy."""
    from pairnet_b200 import _native as nat
    lib = nat.load()
    x, wt, bias = _t((B, 256, H, W), 68), _t((256, 256, 1, 1), 69, 0.1), _t((256,), 70)
    ref = F.conv2d(x.double(), wt.double(), bias.double())
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    wc, bc = wt.cuda(), bias.cuda()
    need = lib.pn_conv1x1_nhwc_to_nchw_workspace_bytes(256)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    y = torch.empty((B, 256, H, W), device="cuda")
    nat.check(lib.pn_conv1x1_nhwc_to_nchw(xc.data_ptr(), wc.data_ptr(), bc.data_ptr(), y.data_ptr(), B, H * W, 256,
                                          ws.data_ptr(), need, torch.cuda.current_stream().cuda_stream), "conv1x1")
    assert rel_err(y.cpu(), ref) < 5e-6


def test_pixel_decoder_native_tail_vs_torch():
    """FPN tail on the CUDA library vs the cuDNN/ATen restatement (same device, same encoder output)."""
    m = _pixel_decoder()
    feats = [_t((2, 256, 40, 56), 41).cuda().contiguous(memory_format=torch.channels_last),
             _t((2, 512, 20, 28), 42).cuda().contiguous(memory_format=torch.channels_last),
             _t((2, 1024, 10, 14), 43).cuda().contiguous(memory_format=torch.channels_last),
             _t((2, 2048, 5, 7), 44).cuda().contiguous(memory_format=torch.channels_last)]
    with torch.no_grad():
        m.tail_impl = "native"
        mf_n, mem_n = m(feats)
        m.tail_impl = "torch"
        mf_t, mem_t = m(feats)
    # token-major (channels_last) for the head's tensor-core mask path; NCHW-contiguous when that path is off
    assert mf_n.is_contiguous(memory_format=torch.channels_last) and mf_n.shape == mf_t.shape
    for a, b in zip(mem_n, mem_t):
        assert torch.equal(a, b)
    assert rel_err(mf_n, mf_t) < 1e-3  # the 3x3 output conv (cuDNN, TF32) is shared; the 1x1 is fp32-accurate here


@pytest.mark.parametrize("cin,H,W,bias", [(2048, 25, 42, True), (1024, 50, 84, True), (512, 100, 167, True),
                                          (256, 200, 334, False), (1536, 32, 32, True), (192, 64, 64, False)])
def test_pixel_decoder_conv1x1_native_vs_cudnn_fp32(cin, H, W, bias):
    """1x1 input / lateral convolutions of the pixel decoder on the tcgen05 GEMM (3xTF32) vs cuDNN with TF32 disabled
    (true fp32), at the R50 (800x1333) and Swin-L (1024^2) channel counts; GroupNorm follows on the native kernel."""
    from pairnet_b200.upstream.pixel_decoder import ConvModule
    torch.manual_seed(cin)
    m = ConvModule(cin, 256, 1, bias=bias).cuda().eval()
    with torch.no_grad():
        m.gn.weight.uniform_(0.5, 1.5)
        m.gn.bias.normal_(0, 0.1)
    x = _t((2, cin, H, W), 11 + cin).cuda().contiguous(memory_format=torch.channels_last)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            y_conv = m._native_conv1x1(x)
            assert y_conv is not None and y_conv.is_contiguous(memory_format=torch.channels_last)
            ref_conv = m.conv(x)
            got = m(x)
            ConvModule.native_conv1x1 = False
            try:
                ref = m(x)
            finally:
                ConvModule.native_conv1x1 = True
    finally:
        torch.backends.cudnn.allow_tf32 = old
    # 3xTF32: the tensor pipe accumulates with truncation, ~1e-5 of the scale at K = 2048 (DESIGN lesson 2)
    assert rel_err(y_conv, ref_conv) < 3e-5
    assert rel_err(got, ref) < 5e-5
    # PyTorch's default (cudnn.allow_tf32 = True): one TF32 pass, like the cuDNN convolution it replaces
    with torch.no_grad():
        y_tf32, ref_tf32 = m._native_conv1x1(x), m.conv(x)
    assert rel_err(y_tf32, ref_tf32) < 2e-3 and rel_err(y_tf32, ref_conv) < 2e-3
    # a layout the GEMM does not take (NCHW-contiguous) is declined, not silently mis-read
    assert m._native_conv1x1(x.contiguous()) is None or cin == 1


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 50, 67), (1, 64, 400, 667), (1, 8, 5, 4)])
def test_maxpool3x3s2_nhwc_bit_exact(B, C, H, W):
    from pairnet_b200 import _native as nat
    x = _t((B, C, H, W), 71).cuda().contiguous(memory_format=torch.channels_last)
    ref = F.max_pool2d(x, 3, 2, 1)
    y = torch.empty_like(ref, memory_format=torch.channels_last)
    nat.check(nat.load().pn_maxpool3x3s2_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, C,
                                              torch.cuda.current_stream().cuda_stream), "maxpool")
    assert torch.equal(y, ref)


def test_backbone_fused_epilogue_path_matches_module_path():
    """upstream plumbing: BN-folded channels_last ResNet-50 with cuDNN fused epilogues, the native stem max-pool and the
    projection-shortcut bias folded into conv3's epilogue vs the plain module path (both TF32 convs on the GPU)."""
    from pairnet_b200.upstream.backbone import ResNet
    torch.manual_seed(3)
    m = ResNet(depth=50, norm_eval=True).cuda().eval()
    with torch.no_grad():
        for mod in m.modules():  # non-trivial frozen-BN statistics
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.1)
        x = torch.randn(2, 3, 128, 160, device="cuda")
        fast = m(x)
        type(m).fuse_epilogues = False
        try:
            slow = m(x)
        finally:
            type(m).fuse_epilogues = True
    for a, b in zip(fast, slow):
        assert a.shape == b.shape
        assert rel_err(a, b) < 5e-3  # TF32 convolutions with different algorithms / summation orders


# --------------------------------------------------------------------------- PPN at scale (BASELINE config 5)
@pytest.mark.parametrize("B,N", [(12, 100), (7, 200), (3, 400), (30, 37), (9, 128), (8, 130), (600, 100), (5, 256),
                                 (4, 388), (20, 64), (40, 48), (150, 104), (3, 512)])
def test_pair_matrix_tcgen05_and_topk_batched(B, N):
    """Pair matrix on the tcgen05 kernel (both operands split in the SM, 3-D TMA with OOB zero fill) + batched
    top-k: importance to fp32 accuracy of the fp64 product, indices bit-exact for the matrix the kernel wrote."""
    import torch.nn.functional as F
    from pairnet_b200 import _native as nat, ops
    K = 100
    g = torch.Generator().manual_seed(1000 + N)
    s = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1)
    o = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1)
    plan = ops.PpnPlan(B, N, K, "cuda")
    assert B * N >= 1024 and nat.load().pn_get_option(7) == 1  # -> tcgen05 path
    imp, idx, sp, op = plan.run_embeds(s.cuda(), o.cuda())
    torch.cuda.synchronize()
    ref = torch.matmul(s.double(), o.double().transpose(1, 2))
    assert float((imp.cpu().double() - ref).abs().max()) < 2e-6  # |cos| <= 1: absolute = relative to the scale
    from oracle.head import stable_topk
    rv, ri = torch.topk(imp.cpu().flatten(1), K)
    for b in (idx.cpu() != ri).any(1).nonzero().flatten().tolist():  # exact fp32 ties: torch's order is unspecified
        ri[b] = torch.from_numpy(stable_topk(imp[b].flatten().cpu().numpy(), K))
    assert torch.equal(idx.cpu(), ri)
    assert torch.equal(sp.cpu(), torch.div(ri, N, rounding_mode="trunc"))
    assert torch.equal(op.cpu(), torch.remainder(ri, N))
    # the exact-FFMA pair matrix agrees to fp32 noise
    nat.load().pn_set_option(7, 0)
    try:
        imp2 = plan.run_embeds(s.cuda(), o.cuda())[0].clone()
    finally:
        nat.load().pn_set_option(7, 1)
    assert float((imp2 - imp).abs().max()) < 2e-6
    # fused pair-matrix + top-k kernel (default where supported) == pair-matrix kernel followed by the top-k kernel
    nat.load().pn_set_option(nat.PN_OPT_PPN_FUSED_TOPK, 0)
    try:
        imp3, idx3, sp3, op3 = [t.clone() for t in plan.run_embeds(s.cuda(), o.cuda())]
    finally:
        nat.load().pn_set_option(nat.PN_OPT_PPN_FUSED_TOPK, 1)
    assert float((imp3 - imp).abs().max()) < 2e-6
    r3 = torch.topk(imp3.cpu().flatten(1), K).indices
    for b in (idx3.cpu() != r3).any(1).nonzero().flatten().tolist():
        r3[b] = torch.from_numpy(stable_topk(imp3[b].flatten().cpu().numpy(), K))
    assert torch.equal(idx3.cpu(), r3) and torch.equal(sp3.cpu() * N + op3.cpu(), r3)


@pytest.mark.parametrize("N", [100, 200, 400])
def test_pair_topk_fused_adversarial_images_take_the_exact_path(N):
    """Images the fused kernel cannot threshold (constant / heavily tied / clustered matrices overflow its candidate
    buffer) are flagged and redone by the exact kernel; results stay bit-exact, ties by ascending flat index."""
    import torch.nn.functional as F
    from oracle.head import stable_topk
    from pairnet_b200 import ops
    B, K = 16, 100
    g = torch.Generator().manual_seed(77 + N)
    s = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1)
    o = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1)
    s[1] = s[1, :1]                      # constant rows: every column of the matrix repeats N times
    o[1] = o[1, :1]                      # ... and every entry is the same value
    s[4, :, :] = s[4, :1]                # rank-one: N-fold ties
    o[7, N // 2:] = o[7, :N - N // 2]    # duplicated objects: pairwise ties
    o[12] = -o[12, :1]                   # constant columns, negative values
    imp, idx, sp, op = ops.PpnPlan(B, N, K, "cuda").run_embeds(s.cuda(), o.cuda())
    torch.cuda.synchronize()
    ref = torch.matmul(s.double(), o.double().transpose(1, 2))
    assert float((imp.cpu().double() - ref).abs().max()) < 2e-6
    for b in range(B):
        want = torch.from_numpy(stable_topk(imp[b].flatten().cpu().numpy(), K))
        assert torch.equal(idx[b].cpu(), want), b
        assert torch.equal(sp[b].cpu() * N + op[b].cpu(), want), b


@pytest.mark.parametrize("B,N", [(12, 100), (7, 200), (3, 400), (30, 52), (9, 128), (8, 132), (600, 100), (5, 256),
                                 (4, 388), (1, 100), (2, 100), (150, 104), (3, 512)])
def test_pair_topk_bf16(B, N):
    """bf16 entry point (`pn_ppn_pair_topk_bf16`, SURVEY 8b / config 5 "bf16 run"): bf16 x bf16 products are exact in
    fp32, so the matrix equals the fp64 product of the bf16-rounded embeddings to fp32 summation noise; indices are
    bit-exact for the matrix the kernel wrote (ties by ascending flat index)."""
    import torch.nn.functional as F
    from oracle.head import stable_topk
    from pairnet_b200 import ops
    K = 100
    g = torch.Generator().manual_seed(2000 + N)
    s = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1).to(torch.bfloat16)
    o = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1).to(torch.bfloat16)
    if B >= 8:  # adversarial images: constant matrix, N-fold ties, duplicated columns -> flagged, exact kernel
        s[1] = s[1, :1]
        o[1] = o[1, :1]
        s[4, :, :] = s[4, :1]
        o[6, N // 2:] = o[6, :N - N // 2]
    plan = ops.PpnPlan(B, N, K, "cuda")
    imp, idx, sp, op = plan.run_embeds_bf16(s.cuda(), o.cuda())
    torch.cuda.synchronize()
    ref = torch.matmul(s.double(), o.double().transpose(1, 2))
    assert float((imp.cpu().double() - ref).abs().max()) < 1e-6
    for b in range(B):
        want = torch.from_numpy(stable_topk(imp[b].flatten().cpu().numpy(), K))
        assert torch.equal(idx[b].cpu(), want), b
        assert torch.equal(sp[b].cpu() * N + op[b].cpu(), want), b
    # the fp32 entry point on the same (bf16-representable) values writes the same matrix to fp32 noise
    imp32 = ops.PpnPlan(B, N, K, "cuda").run_embeds(s.float().cuda(), o.float().cuda())[0]
    assert float((imp32 - imp).abs().max()) < 2e-6


def test_pair_topk_bf16_rejects_fp32_and_bad_shapes():
    from pairnet_b200 import _native as nat, ops
    plan = ops.PpnPlan(4, 100, 100, "cuda")
    with pytest.raises(TypeError):
        plan.run_embeds_bf16(torch.zeros(4, 100, 256, device="cuda"), torch.zeros(4, 100, 256, device="cuda"))
    bad = ops.PpnPlan(4, 101, 100, "cuda")   # N % 4 != 0: no silent fallback
    z = torch.zeros(4, 101, 256, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(nat.NativeError):
        bad.run_embeds_bf16(z, z)


@pytest.mark.parametrize("bf16", [False, True])
def test_pair_topk_speculative_threshold_failures_are_redone_exactly(bf16):
    """PN_OPT_PPN_SPECULATE: a CTA thresholds image i+1 with the ~1.5 K-th value of image i.  Consecutive images of a CTA
    (b, b + 148, b + 296 on a 148-SM B200) are given scales 1, 0.01 and 100 here, so the guess leaves no candidate for the
    second and overflows the candidate buffer for the third: both are flagged and redone by the exact kernel; every
    image's indices stay exact (ties by ascending flat index)."""
    import torch.nn.functional as F
    from oracle.head import stable_topk
    from pairnet_b200 import _native as nat, ops
    B, N, K = 3 * 148 + 5, 100, 100
    g = torch.Generator().manual_seed(9)
    s = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1)
    o = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1)
    s[148:296] *= 0.01
    s[296:444] *= 100.0
    if bf16:
        s, o = s.to(torch.bfloat16), o.to(torch.bfloat16)
    plan = ops.PpnPlan(B, N, K, "cuda")
    assert nat.load().pn_get_option(nat.PN_OPT_PPN_SPECULATE) == 1
    run = plan.run_embeds_bf16 if bf16 else plan.run_embeds
    imp, idx, sp, op = run(s.cuda(), o.cuda())
    torch.cuda.synchronize()
    ref = torch.matmul(s.double(), o.double().transpose(1, 2))
    scale = torch.ones(B, 1, 1, dtype=torch.float64)      # |s_i| |o_j| of each image: the scale 3xTF32 errors are relative to
    scale[148:296], scale[296:444] = 0.01, 100.0
    assert float(((imp.cpu().double() - ref).abs() / scale).max()) < 2e-6
    for b in list(range(0, B, 37)) + [147, 148, 149, 295, 296, 297, 443, 444, B - 1]:
        want = torch.from_numpy(stable_topk(imp[b].flatten().cpu().numpy(), K))
        assert torch.equal(idx[b].cpu(), want), b
        assert torch.equal(sp[b].cpu() * N + op[b].cpu(), want), b


def test_ppn_l2_chunked_batch_equals_unchunked():
    """Batches whose pair matrices exceed the L2 chunk budget are walked chunk by chunk: same results."""
    import torch.nn.functional as F
    from pairnet_b200 import ops
    B, N, K = 1400, 100, 100   # 56 MB of pair matrices > 48 MB chunk budget
    g = torch.Generator().manual_seed(5)
    s = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1).cuda()
    o = F.normalize(torch.randn(B, N, 256, generator=g), dim=-1).cuda()
    imp, idx, sp, op = ops.PpnPlan(B, N, K, "cuda").run_embeds(s, o)
    ref = torch.matmul(s, o.transpose(1, 2))
    assert float((imp - ref).abs().max()) < 5e-6
    from oracle.head import stable_topk
    ri = torch.topk(imp.flatten(1), K).indices
    for b in (idx != ri).any(1).nonzero().flatten().tolist():  # exact fp32 ties: torch's order is unspecified
        ri[b] = torch.from_numpy(stable_topk(imp[b].flatten().cpu().numpy(), K)).cuda()
    assert torch.equal(idx, ri)
    assert torch.equal(sp * N + op, ri)


@pytest.mark.parametrize("N,K", [(100, 100), (100, 300), (200, 100), (400, 100), (30, 900), (16, 100)])
def test_topk_threshold_path_equals_radix_path(N, K):
    """Default top-k (local-maxima threshold + candidate ranking) == forced exact radix select, incl. K larger
    than the CTA (falls back), tiny matrices, smooth (spatially correlated) and heavily tied inputs."""
    from oracle.head import stable_topk
    from pairnet_b200 import _native as nat, ops
    rng = np.random.default_rng(N * 7 + K)
    smooth = torch.from_numpy(rng.standard_normal((1, N, N)).astype(np.float32))
    smooth = torch.nn.functional.avg_pool2d(smooth[None], 7, 1, 3)[0]
    ramp = torch.arange(N * N, dtype=torch.float32).view(1, N, N) / (N * N)        # ascending: adversarial for blocks
    tied = torch.from_numpy(rng.integers(0, 3, size=(1, N, N)).astype(np.float32))
    imp = torch.cat([smooth, ramp, -ramp, tied, _t((1, N, N), 99)]).cuda()
    a = ops.topk_pairs(imp, K)
    nat.load().pn_set_option(6, 1)
    try:
        b = ops.topk_pairs(imp, K)
    finally:
        nat.load().pn_set_option(6, 0)
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    for i in range(imp.shape[0]):
        assert a[0][i].cpu().tolist() == stable_topk(imp[i].flatten().cpu().numpy(), K).tolist()
