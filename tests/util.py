"""Shared helpers for the parity tests: identical synthetic weights in the CPU oracle and in the
B200 head, tolerance helpers.  (Tests are the only place allowed to import ``oracle``.)"""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
RTOL_LOGITS = 1e-3  # BASELINE.json north_star: "within 1e-3 relative for fp32 logits"


def rel_err(a, b):
    """max|a-b| / max|b|  (the 'relative' of the 1e-3 logits bar: normalised by the tensor's scale)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def oracle_small_head(seed=10086, dtype=torch.float32):
    from oracle.make_golden import build_small_head
    return build_small_head(seed, dtype)


def product_head_cfg():
    from pairnet_b200.registry import Config
    cfg = Config.fromfile(os.path.join(ROOT, "configs", "pairnet_r50_b200.py"), import_custom_modules=False)
    return cfg.model.bbox_head


def product_small_head(oracle_head, device="cuda"):
    """B200 CrossHead2 (no pixel decoder) carrying exactly the oracle head's weights."""
    from pairnet_b200.registry import build_head
    cfg = product_head_cfg()
    cfg["pixel_decoder"] = None
    head = build_head(cfg)
    missing = head.load_state_dict(oracle_head.state_dict(), strict=True)
    return head.to(device).eval()


def check_topk_tie_aware(importance, sub_pos, obj_pos, ref_sub, ref_obj, tol):
    """Selected pairs must equal the oracle's, except where the oracle's own values are within `tol`
    (near-ties can legitimately swap under a different fp32 summation order)."""
    imp = torch.as_tensor(importance).double().cpu()
    B, N, _ = imp.shape
    n_swapped = 0
    for b in range(B):
        mine = (torch.as_tensor(sub_pos[b]).cpu() * N + torch.as_tensor(obj_pos[b]).cpu()).tolist()
        ref = (torch.as_tensor(ref_sub[b]).cpu() * N + torch.as_tensor(ref_obj[b]).cpu()).tolist()
        flat = imp[b].flatten()
        for r, (i, j) in enumerate(zip(mine, ref)):
            if i != j:
                n_swapped += 1
                assert abs(float(flat[i]) - float(flat[j])) <= tol, (
                    f"image {b} rank {r}: picked {i} ({float(flat[i])}) vs oracle {j} ({float(flat[j])})")
        assert len(set(mine)) == len(mine)
    return n_swapped
