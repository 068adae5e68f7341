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
from . import ops
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
    def __init__(self, cfg, device, total_steps=100000, dtype=torch.bfloat16, ddp=None, seed=0, cuda_graph=False):
        """cuda_graph=True (single process, CUDA): after a few eager steps the whole step — forward, losses, backward,
        gradient clipping and AdamW — is captured once into a CUDA graph and replayed from static input buffers; the host
        then issues one graph launch per step instead of ~2000 kernel launches.  Dropout stays random: torch's own dropout
        through the graph-registered generator, the in-kernel masks through the device step counter registered with
        sv_dropout_seed_offset."""
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
        self.use_ddp = ddp if ddp is not None else (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
        # cuda_graph with several ranks: no DDP wrapper — graph A (forward + losses + backward into one flat gradient
        # buffer), ONE eager NCCL all-reduce (mean) of that buffer, graph B (clipping + AdamW).  No collective is captured;
        # the contrastive exchange inside graph A is the native peer-memory kernel with a device-resident epoch.
        self.graph_mode = bool(cuda_graph) and self.device.type == "cuda"
        self.dp_graph = self.graph_mode and self.use_ddp
        if self.dp_graph:
            self.use_ddp = False
        warm = self.cfg.solver.sched.args.warmup_steps * self.cfg.num_gpu
        mr = self.cfg.solver.sched.args.get("minimum_ratio", 1e-5)
        self._lr_lambda = lambda s: M.warmup_cosine(s, warm, total_steps, minimum_ratio=mr)
        if self.graph_mode:
            # capturable AdamW reads lr from a device tensor; the warm-up/cosine schedule fills it from the host per step
            self._base_lrs = [float(g.get('lr', self.cfg.solver.lr)) for g in groups]
            for g in groups:
                g['lr'] = torch.tensor(float(g.get('lr', self.cfg.solver.lr)) * self._lr_lambda(0), device=self.device)
            self.optimizer = torch.optim.AdamW(groups, fused=True, capturable=True, **kw)
            self.scheduler = None
            self._sched_step = 0
            self._step_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
            from . import _lib
            _lib.check(_lib.gps(), _lib.gps().sv_dropout_seed_offset(self._step_counter.data_ptr()), "sv_dropout_seed_offset")
            self.graph = self.graph_opt = self.static_batch = self.static_loss = self.flat_grads = None
        else:
            self.optimizer = torch.optim.AdamW(groups, lr=self.cfg.solver.lr, fused=self.device.type == "cuda", **kw)
            self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, self._lr_lambda)
        self.grad_norm = self.cfg.solver.get("grad_norm")
        self.ddp = None
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

    def close(self):
        """Unregister the process-wide hooks of graph mode (the dropout seed counter lives in this object's memory, the
        bf16 weight shadows reference its parameters)."""
        if getattr(self, "_step_counter", None) is not None:
            self._step_counter = None
            try:
                from . import _lib
                _lib.gps().sv_dropout_seed_offset(None)
                ops.clear_shadows()
            except Exception:      # interpreter shutdown: modules may already be gone
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- CUDA-graph path ---------------------------------------------------------------------------------------------
    def _raw_fwd_bwd(self):
        self._step_counter.add_(1)
        ops.refresh_shadows()      # one multi-tensor fp32 -> bf16 copy of all weights (instead of a cast per linear)
        if self.flat_grads is not None:
            self.flat_grads.zero()
        else:
            self.optimizer.zero_grad(set_to_none=False)
        # gradients live in the flat buffer: the wgrad kernels accumulate into it directly (no AccumulateGrad adds)
        prev, ops.DIRECT_GRAD[0] = ops.DIRECT_GRAD[0], self.flat_grads is not None
        try:
            with torch.autocast(self.device.type, dtype=self.dtype, enabled=self.dtype != torch.float32):
                total, _ = self.module(dict(self.static_batch))
            total.backward()
        finally:
            ops.DIRECT_GRAD[0] = prev
        return total.detach()

    def _raw_opt(self):
        if self.grad_norm is not None:
            if self.flat_grads is not None:
                self.flat_grads.clip_norm_(self.grad_norm)     # same math as clip_grad_norm_, two kernels on the flat buffer
            else:
                torch.nn.utils.clip_grad_norm_(self.parameters(), self.grad_norm, foreach=True)
        self.optimizer.step()

    def _raw_step(self):
        loss = self._raw_fwd_bwd()
        if self.flat_grads is not None:
            self.flat_grads.all_reduce_mean()
        self._raw_opt()
        return loss

    def _advance_lr(self):
        self._sched_step += 1
        f = self._lr_lambda(self._sched_step)
        for g, base in zip(self.optimizer.param_groups, self._base_lrs):
            g['lr'].fill_(base * f)

    def _capture(self, data_dict):
        self.static_batch = {k: v.clone() for k, v in data_dict.items() if torch.is_tensor(v)}
        if self.dtype == torch.bfloat16:
            ops.clear_shadows()
            ops.register_shadows(self.module)
        if self.dp_graph:
            sync_module_state(self.module)
        self.flat_grads = FlatGrads(self.parameters())
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(3):          # real optimisation steps: allocates grads / optimizer state outside the graph pool
                self._raw_step()
                self._advance_lr()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        from . import _lib
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        if not self.dp_graph:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._raw_step()
        else:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._raw_fwd_bwd()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
                self._raw_opt()
        self.native_launches_per_step = _lib.launch_count() - n0   # native kernel nodes replayed per step

    def _dp_graph_possible(self):
        """Graph A must not contain an NCCL collective: with several ranks the contrastive exchange has to be the native
        peer-memory kernel.  If symmetric memory could not be set up (losses then use NCCL all_gather), run eager DDP."""
        from . import fused_gather
        return all(v is not None for v in fused_gather._state.values())

    def _graph_step(self, data_dict):
        if not self.probed:
            self._probe_unused(data_dict)
            if self.dp_graph and not self._dp_graph_possible():
                self.graph_mode = self.dp_graph = False
                self.use_ddp = True
                self.ddp = nn.parallel.DistributedDataParallel(
                    self.module, device_ids=[self.device.index], find_unused_parameters=False,
                    gradient_as_bucket_view=True, bucket_cap_mb=64)
                return self.step(data_dict)
        if self.graph is None:
            self._capture(data_dict)
        for k, dst in self.static_batch.items():
            src = data_dict[k]
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if self.dp_graph:
            self.flat_grads.all_reduce_mean()
            self.graph_opt.replay()
        self._advance_lr()
        return self.static_loss

    def step(self, data_dict):
        """data_dict: tensors already on self.device. Returns the (detached) total loss tensor — no host sync."""
        if self.graph_mode:
            return self._graph_step(data_dict)
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
        if self.scheduler is not None:
            self.scheduler.step()
        else:                       # capturable optimizer (graph mode was requested but is not possible): tensor lr
            self._advance_lr()
        return total.detach()


class FlatGrads:
    """Gradients of `params` as views of ONE flat fp32 buffer, so that the data-parallel mean is a single NCCL call (the
    reference's DDP reduces the same 491 MB in 25 MB buckets, trainer/build.py:66-75)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=self.params[0].dtype, device=self.params[0].device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def clip_norm_(self, max_norm, eps=1e-6):
        """torch.nn.utils.clip_grad_norm_ (L2) on the flat buffer: the norm of the per-tensor norms is the norm of the
        concatenation."""
        total = torch.linalg.vector_norm(self.flat, 2.0)
        self.flat.mul_((max_norm / (total + eps)).clamp(max=1.0))
        return total

    def all_reduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))


def sync_module_state(module, src=0):
    """What DDP does at construction: every rank starts from rank `src`'s parameters and buffers."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)


def batch_to_device(np_batch, device, pinned=None, non_blocking=True):
    """numpy data_dict (synthetic.scene_batch) -> tensors on device; `pinned` = matching dict of pinned host tensors."""
    out = {}
    for k, v in np_batch.items():
        src = pinned[k] if pinned is not None else torch.from_numpy(v)
        out[k] = src.to(device, non_blocking=non_blocking)
    return out
