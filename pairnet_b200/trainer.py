"""One data-parallel training step of Pair-Net (BASELINE config 3).

    forward (backbone + pixel decoder: CUDA / cuDNN, no grad;  head: torch_head.py, autograd)
    -> targets + losses (training.py)  -> backward
    -> gradient ALL-REDUCE over NCCL (bucketed, launched from autograd hooks while backward is still running)
    -> grad clip (max_norm 0.1) -> AdamW (lr 1e-4, wd 1e-4, lr_mult 0.1 on the Mask2Former decoder, no decay on norms)

Reference recipe: ``configs/mask2former/pairnet.py:352-372`` (optimizer, paramwise_cfg, grad_clip), ``tools/train.py``
(mmcv EpochBasedRunner + MMDistributedDataParallel with ``find_unused_parameters=True``).  Here the set of parameters
that receive gradients is static (``torch_head.trainable_parameters``), so nothing has to be discovered per step and
the buckets are fixed: the only collective of the whole system is this all-reduce (the forward shards by image)."""
import torch
import torch.distributed as dist

from . import torch_head as th


class GradReducer:
    """Bucketed gradient all-reduce.  Parameters are packed (in reverse registration order = roughly the order their
    gradients become ready) into flat fp32 buckets; ``p.grad`` is a VIEW into its bucket, so autograd accumulates
    straight into the communication buffer.  A post-accumulate hook per parameter counts its bucket down and launches
    ``all_reduce(async_op=True)`` on the bucket the moment it is complete -- NCCL's stream runs it while the autograd
    engine keeps producing the earlier layers' gradients.  ``finish()`` waits and averages."""

    def __init__(self, params, bucket_bytes=25 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # [flat tensor, [params], pending count]
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._close(cur)
        self.bytes = sum(b[0].numel() * 4 for b in self.buckets)
        self._works = []
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def _close(self, ps):
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=ps[0].device)
        off = 0
        idx = len(self.buckets)
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)
            p._pn_bucket = idx
            off += p.numel()
        self.buckets.append([flat, list(ps), len(ps)])

    def _hook(self, p):
        b = self.buckets[p._pn_bucket]
        b[2] -= 1
        if b[2] == 0 and self.world > 1:
            self._works.append(dist.all_reduce(b[0], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Wait for the in-flight buckets, average, re-arm the counters.  Returns the number of collectives issued."""
        n = len(self._works)
        for w in self._works:
            w.wait()
        self._works = []
        for b in self.buckets:
            if b[2] != 0:   # a parameter of this bucket got no gradient this step: reduce what is there, late
                if self.world > 1:
                    dist.all_reduce(b[0], op=dist.ReduceOp.SUM, group=self.group)
                    n += 1
            if self.world > 1:
                b[0].div_(self.world)
            b[2] = len(b[1])
        return n

    def zero_grad(self):
        for b in self.buckets:
            b[0].zero_()

    def close(self):
        for h in self._handles:
            h.remove()


def param_groups(named_params, lr, weight_decay):
    """mmcv DefaultOptimizerConstructor with the reference's paramwise_cfg (configs/mask2former/pairnet.py:356-366)."""
    groups = {}
    for n, p in named_params:
        lr_mult = 0.1 if any(k in n for k in ("backbone", "transformer_decoder", "pixel_decoder", "decoder_input_projs")) else 1.0
        is_norm = ".norms." in n or "post_norm" in n
        key = (lr_mult, is_norm)
        groups.setdefault(key, []).append(p)
    return [dict(params=ps, lr=lr * m, weight_decay=0.0 if is_norm else weight_decay) for (m, is_norm), ps in groups.items()]


class TrainStep:
    def __init__(self, model, scope="head", lr=1e-4, weight_decay=1e-4, max_norm=0.1, bucket_bytes=25 << 20,
                 amp_dtype=None):
        """amp_dtype = torch.bfloat16: BASELINE config 3's bf16 training -- the differentiable head (torch ops) runs under
        autocast, master weights, gradients (the all-reduced buckets) and AdamW state stay fp32."""
        self.model, self.scope, self.max_norm = model, scope, max_norm
        self.amp_dtype = amp_dtype
        head = model.bbox_head
        head.train_scope = scope
        named = th.trainable_parameters(head, scope)
        keep = {id(p) for _, p in named}
        for p in model.parameters():
            p.requires_grad_(id(p) in keep)
        model.eval()                          # frozen BN statistics (configs/mask2former/pairnet.py:15-16) ...
        head.relation_decoder.train()         # ... but the relation decoder's ffn_drop = 0.1 is live in training
        head.transformer_decoder.train()
        self.params = [p for _, p in named]
        self.reducer = GradReducer(self.params, bucket_bytes)
        fused = self.params[0].is_cuda
        self.opt = torch.optim.AdamW(param_groups(named, lr, weight_decay), fused=fused)
        self.num_params = sum(p.numel() for p in self.params)

    def __call__(self, img, img_metas, gt_rels, gt_labels, gt_masks):
        with torch.autocast(img.device.type, dtype=self.amp_dtype or torch.bfloat16, enabled=self.amp_dtype is not None):
            losses = self.model.forward_train(img, img_metas, gt_rels=gt_rels, gt_bboxes=None, gt_labels=gt_labels,
                                              gt_masks=gt_masks)
        losses = {k: v.float() for k, v in losses.items()}
        total = sum(losses.values())          # mmdet _parse_losses: every key containing "loss"
        total.backward()
        self.collectives = self.reducer.finish()
        if self.max_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.max_norm)
        self.opt.step()
        self.reducer.zero_grad()              # gradients live in the buckets: zeroed in place, views stay valid
        return losses


def synthetic_targets(batch, hw, seed, device, num_gt=12, num_rel=10, num_object_classes=133, num_relations=56):
    """per image 12 rectangle masks, labels U{0..132}, 10 triplets [sub, obj, predicate]
    (sub != obj, predicate U{1..56})."""
    g = torch.Generator().manual_seed(seed)
    H, W = hw
    rels, labels, masks = [], [], []
    for _ in range(batch):
        m = torch.zeros(num_gt, H, W, dtype=torch.uint8)
        for k in range(num_gt):
            h = int(torch.randint(max(2, H // 8), max(3, H // 2), (1,), generator=g))
            w = int(torch.randint(max(2, W // 8), max(3, W // 2), (1,), generator=g))
            y0 = int(torch.randint(0, H - h + 1, (1,), generator=g))
            x0 = int(torch.randint(0, W - w + 1, (1,), generator=g))
            m[k, y0:y0 + h, x0:x0 + w] = 1
        masks.append(m.to(device))
        labels.append(torch.randint(0, num_object_classes, (num_gt,), generator=g).to(device))
        s = torch.randint(0, num_gt, (num_rel,), generator=g)
        o = (s + torch.randint(1, num_gt, (num_rel,), generator=g)) % num_gt
        p = torch.randint(1, num_relations + 1, (num_rel,), generator=g)
        rels.append(torch.stack([s, o, p], 1).to(device))
    return rels, labels, masks
