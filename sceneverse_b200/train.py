"""One GPS pre-training step (reference: trainer/openvocab_trainer.py:18-46 `train_step` + `backward`,
trainer/build.py:66-75,121,135-145): forward -> list loss -> backward -> clip_grad_norm_(5.0) -> AdamW -> warm-up/cosine.

Data parallelism: one process per GPU, identical replicas, NCCL bucketed gradient all-reduce overlapped
with backward (torch DDP over NVLink) and the embedding all-gather inside the contrastive losses.  Parameters that
never receive a gradient on this path (the reference needs find_unused_parameters=True for them,
trainer/build.py:66) are listed statically and frozen instead of being searched for every step.
"""
import torch
import torch.distributed as dist
from torch import nn

from . import model as M
from .modules import losses as L

# parameters of the reference model that are constructed but never used in OpenVocab.forward
STATIC_UNUSED = ("point_encoder.sem_cls_embed_layer.", "point_encoder.sem_mask_embeddings.", "lang_encoder.model.pooler.")


class StepModule(nn.Module):
    """model + loss in one module so that a single DDP wrapper covers every trainable parameter (incl. logit_scale)."""

    def __init__(self, cfg):
        super().__init__()
        self.model = M.OpenVocab(cfg)
        c = M.to_cfg(cfg)
        self.loss = L.Loss(c.model.loss_list, c.model.vis_loss_list, num_gpu=c.num_gpu)
        for n, p in self.model.named_parameters():
            if n.startswith(STATIC_UNUSED):
                p.requires_grad = False

    def forward(self, data_dict):
        data_dict = self.model(data_dict)
        total, all_losses = self.loss(data_dict)
        return total, all_losses


class PretrainStep:
    def __init__(self, cfg, device, total_steps=100000, dtype=torch.bfloat16, ddp=None, seed=0):
        torch.manual_seed(seed)
        self.cfg = M.to_cfg(cfg)
        self.device = torch.device(device)
        self.module = StepModule(cfg).to(self.device)
        self.dtype = dtype
        groups = self.module.model.get_opt_params()
        groups.append({'params': [p for p in self.module.loss.parameters() if p.requires_grad], 'weight_decay': 0.0,
                       'lr': self.cfg.solver.lr})
        kw = dict(self.cfg.solver.optim.args)
        kw["betas"] = tuple(kw.get("betas", (0.9, 0.999)))
        self.optimizer = torch.optim.AdamW(groups, lr=self.cfg.solver.lr, fused=self.device.type == "cuda", **kw)
        warm = self.cfg.solver.sched.args.warmup_steps * self.cfg.num_gpu
        mr = self.cfg.solver.sched.args.get("minimum_ratio", 1e-5)
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(
            self.optimizer, lambda s: M.warmup_cosine(s, warm, total_steps, minimum_ratio=mr))
        self.grad_norm = self.cfg.solver.get("grad_norm")
        self.ddp = None
        self.use_ddp = ddp if ddp is not None else (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
        self.probed = False
        self.module.train()

    def _probe_unused(self, data_dict):
        """One local forward/backward to find the parameters this configuration never touches (e.g. the object LM head
        when `obj_cls_post_logits` is in no loss, all_pretrain.yaml:246-258).  They are frozen once, so DDP can run with
        find_unused_parameters=False instead of the reference's per-step graph search (trainer/build.py:66)."""
        self.optimizer.zero_grad(set_to_none=True)
        with torch.autocast(self.device.type, dtype=self.dtype, enabled=self.dtype != torch.float32):
            total, _ = self.module(dict(data_dict))
        total.backward()
        unused = [n for n, p in self.module.named_parameters() if p.requires_grad and p.grad is None]
        for n, p in self.module.named_parameters():
            if p.requires_grad and p.grad is None:
                p.requires_grad = False
        self.optimizer.zero_grad(set_to_none=True)
        for g in self.optimizer.param_groups:
            g['params'] = [p for p in g['params'] if p.requires_grad]
        self.unused_parameters = unused
        if self.use_ddp:
            self.ddp = nn.parallel.DistributedDataParallel(
                self.module, device_ids=[self.device.index] if self.device.type == "cuda" else None,
                find_unused_parameters=False, gradient_as_bucket_view=True, bucket_cap_mb=64)
        self.probed = True

    def parameters(self):
        return [p for p in self.module.parameters() if p.requires_grad]

    def step(self, data_dict):
        """data_dict: tensors already on self.device. Returns the (detached) total loss tensor — no host sync."""
        if not self.probed:
            self._probe_unused(data_dict)
        net = self.ddp if self.ddp is not None else self.module
        self.optimizer.zero_grad(set_to_none=True)
        with torch.autocast(self.device.type, dtype=self.dtype, enabled=self.dtype != torch.float32):
            total, _ = net(data_dict)
        total.backward()
        if self.grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.parameters(), self.grad_norm, foreach=True)
        self.optimizer.step()
        self.scheduler.step()
        return total.detach()


def batch_to_device(np_batch, device, pinned=None, non_blocking=True):
    """numpy data_dict (synthetic.scene_batch) -> tensors on device; `pinned` = matching dict of pinned host tensors."""
    out = {}
    for k, v in np_batch.items():
        src = pinned[k] if pinned is not None else torch.from_numpy(v)
        out[k] = src.to(device, non_blocking=non_blocking)
    return out
