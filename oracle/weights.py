"""Version-independent synthetic weights / inputs for parity tests (TEST INFRASTRUCTURE).

Weights are drawn from a numpy ``Generator(PCG64(seed))`` in ``state_dict`` order so that the
golden fixtures under ``tests/golden/`` can be regenerated bit-identically on any torch
version (torch's own RNG streams are not guaranteed stable across releases)."""
import numpy as np
import torch


def numpy_state_dict(module, seed=10086, gain=1.0):
    """Fill every float parameter/buffer of ``module`` deterministically; returns the state dict."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = module.state_dict()
    emb = {n + ".weight" if n else "weight" for n, m in module.named_modules() if isinstance(m, torch.nn.Embedding)}
    out = {}
    for name, t in sd.items():
        if not torch.is_floating_point(t):
            out[name] = t.clone()
            continue
        shape = tuple(t.shape)
        leaf = name.rsplit(".", 1)[-1]
        if "running_var" in name:
            a = 1.0 + 0.1 * rng.random(shape)
        elif "running_mean" in name:
            a = 0.05 * rng.standard_normal(shape)
        elif t.dim() >= 2:
            fan_in = int(np.prod(shape[1:]))
            fan_out = shape[0] * int(np.prod(shape[2:])) if t.dim() > 2 else shape[0]
            if name in emb:
                a = rng.standard_normal(shape)  # nn.Embedding default N(0,1)
            else:
                a = gain * np.sqrt(2.0 / (fan_in + fan_out)) * rng.standard_normal(shape)
        elif leaf == "weight":  # norm scale
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        else:  # biases / norm shifts
            a = 0.05 * rng.standard_normal(shape)
        out[name] = torch.from_numpy(np.asarray(a, dtype=np.float64)).to(t.dtype)
    return out


def numpy_tensor(shape, seed, scale=1.0, dtype=torch.float32):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(scale * rng.standard_normal(shape)).to(dtype)


def fixture_state_dict(head, seed=10086):
    """Synthetic head weights used by every fixture / parity test.

    Plain Xavier-scale random weights make the 100 object queries collapse onto one vector after
    9 post-norm layers (every pair then scores the same and top-k degenerates to the conv's
    zero-padding border).  Damping the residual branches and sharpening the Matrix Learner keeps
    the queries distinct so the selected pairs spread over ~50 subject rows -- a harder, more
    realistic test of top-k / gather parity.  Purely a choice of test data."""
    sd = numpy_state_dict(head, seed)
    for k in sd:
        if k.endswith("out_proj.weight") or k.endswith("ffns.0.layers.1.weight"):
            sd[k] = sd[k] * 0.25
        elif "update_importance" in k and k.endswith("weight"):
            sd[k] = sd[k] * 3.0
        elif "update_importance" in k and k.endswith("bias"):
            sd[k] = sd[k] * 0.1
        elif ("sub_query_update" in k or "obj_query_update" in k) and k.endswith("weight"):
            sd[k] = sd[k] * 2.0
    return sd
