"""CPU: the oracle against the committed golden vectors.  ``convtiny_ref_*`` were produced by the REFERENCE's
own ``cnn_factory.ConvTiny``; ``head_ref_*`` by the REFERENCE's own ``CrossHead2.forward`` executed from
``/root/reference`` over constructor shims for the absent mmcv/mmdet (``oracle/pin_reference.py``);
``head_small_*`` are the oracle's own regression anchors."""
import os

import numpy as np
import pytest
import torch

from oracle.head import OConvTiny, stable_topk
from oracle.make_golden import CONV_CASES, HEAD_CASES, build_small_head, small_head_inputs
from oracle.pin_reference import REF_CASES
from oracle.weights import numpy_state_dict, numpy_tensor
from tests.util import GOLDEN, rel_err


@pytest.mark.parametrize("tag,mid,B,N,seed", CONV_CASES)
def test_convtiny_oracle_matches_reference_golden(tag, mid, B, N, seed):
    g = np.load(os.path.join(GOLDEN, f"convtiny_ref_{tag}.npz"))
    m = OConvTiny(mid_channels=mid).eval()
    m.load_state_dict(numpy_state_dict(m, seed))
    x = torch.tanh(numpy_tensor((B, N, N), seed + 100))
    with torch.no_grad():
        y = m(x)
    assert y.shape == (B, N, N)
    assert rel_err(y, g["out"]) < 1e-5


def test_convtiny_oracle_equals_live_reference_if_present():
    ref_path = "/root/reference/pairnet/models/frameworks/cnn_factory.py"
    if not os.path.exists(ref_path):
        pytest.skip("reference tree not present on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_cnn_factory", ref_path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    r = ref.ConvTiny().eval()
    r.load_state_dict(numpy_state_dict(r, 5))
    o = OConvTiny().eval()
    o.load_state_dict(r.state_dict())
    x = torch.tanh(numpy_tensor((1, 50, 50), 6))
    with torch.no_grad():
        assert torch.equal(r(x), o(x))


@pytest.mark.parametrize("tag,B,hw4,seed", HEAD_CASES[1:])
def test_head_oracle_regression(tag, B, hw4, seed):
    g = np.load(os.path.join(GOLDEN, f"head_small_{tag}.npz"))
    head = build_small_head()
    mf, mems = small_head_inputs(B, hw4, seed)
    tr = {}
    with torch.no_grad():
        cls, msk = head.forward_from_memories(mf, mems, trace=tr)
    assert rel_err(cls["rel"], g["rel"]) < 1e-4
    assert rel_err(cls["cls"], g["cls"]) < 1e-4
    assert rel_err(cls["importance"], g["importance"]) < 1e-4
    assert rel_err(msk["mask"][:, :, ::4, ::4], g["mask_sub4"]) < 1e-4
    # indices: equal except for near-ties
    same = (tr["sub_pos"].numpy() == g["sub_pos"]) & (tr["obj_pos"].numpy() == g["obj_pos"])
    assert same.mean() > 0.9


def _ref_case_oracle(B, hw4, seed, N, R):
    from oracle.head import HeadHyper, OCrossHead2
    from oracle.weights import fixture_state_dict
    o = OCrossHead2(HeadHyper(with_pixel_decoder=False, num_obj_query=N, num_rel_query=R)).eval()
    o.load_state_dict(fixture_state_dict(o, 10086))
    mf, mems = small_head_inputs(B, hw4, seed)
    tr = {}
    with torch.no_grad():
        cls, msk = o.forward_from_memories(mf, mems, trace=tr)
    return cls, msk, tr


@pytest.mark.parametrize("tag,B,hw4,seed,N,R", REF_CASES)
def test_head_oracle_matches_reference_forward_golden(tag, B, hw4, seed, N, R):
    """Fixtures minted by the reference's own CrossHead2.forward (bit-equal to the oracle in the dev container;
    here a small tolerance absorbs CPU-kernel differences between hosts)."""
    g = np.load(os.path.join(GOLDEN, f"head_ref_{tag}.npz"))
    cls, msk, tr = _ref_case_oracle(B, hw4, seed, N, R)
    for k in ("cls", "rel", "importance", "sub", "obj"):
        assert cls[k].shape == g[k].shape
        assert rel_err(cls[k], g[k]) < 1e-4, k
    assert rel_err(msk["mask"][:, :, ::4, ::4], g["mask_sub4"]) < 1e-4
    assert rel_err(msk["sub_seg"][:, :, ::4, ::4], g["sub_seg_sub4"]) < 1e-4
    assert rel_err(msk["obj_seg"][:, :, ::4, ::4], g["obj_seg_sub4"]) < 1e-4
    same = (tr["sub_pos"].numpy() == g["sub_pos"]) & (tr["obj_pos"].numpy() == g["obj_pos"])
    assert same.mean() > 0.9


def test_oracle_bit_equals_live_reference_forward_if_present():
    """Dev container only: execute the reference's CrossHead2.forward from /root/reference (subprocess: the
    harness installs mmcv/mmdet constructor shims into sys.modules) and require bit-equality with the oracle."""
    if not os.path.isdir("/root/reference/pairnet"):
        pytest.skip("reference tree not present on this box")
    import subprocess
    import sys
    from tests.util import ROOT
    r = subprocess.run([sys.executable, "-m", "oracle.pin_reference", "--check"], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("pinned ") == len(REF_CASES)


def test_stable_topk_contract():
    v = np.array([1.0, 3.0, 3.0, -1.0, 3.0, 2.0], dtype=np.float32)
    assert stable_topk(v, 4).tolist() == [1, 2, 4, 5]
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1000).astype(np.float32)
    assert stable_topk(x, 50).tolist() == torch.topk(torch.from_numpy(x), 50).indices.tolist()


def test_attach_native_on_the_reference_class_if_present():
    """INTEGRATION.md Option B on the reference's OWN CrossHead2 (built over the oracle's mmcv shims): ``attach_native``
    finds every attribute the native plumbing reads -- the struct fill runs through until the first device requirement
    (CPU parameters -> NativeError), never an AttributeError."""
    if not os.path.isdir("/root/reference/pairnet"):
        pytest.skip("reference tree not present on this box")
    import subprocess
    import sys
    from tests.util import ROOT
    code = (
        "import torch\n"
        "from oracle import pin_reference as pr\n"
        "from pairnet_b200.head import attach_native\n"
        "from pairnet_b200 import _native as nat\n"
        "h = attach_native(pr.build_reference_head())\n"
        "ps = h._hot_params()\n"
        "assert len(ps) > 250 and h._wkey is None and h._pos_cache == {}\n"
        "try:\n"
        "    h.native_weights()\n"
        "    raise SystemExit('expected NativeError on CPU parameters')\n"
        "except nat.NativeError as e:\n"
        "    assert 'CUDA' in str(e)\n"
        "try:\n"
        "    h.forward_from_memories(torch.zeros(1, 256, 8, 8), [torch.zeros(1, 256, 2, 2)] * 3)\n"
        "    raise SystemExit('expected NativeError on CPU tensors')\n"
        "except nat.NativeError:\n"
        "    pass\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
