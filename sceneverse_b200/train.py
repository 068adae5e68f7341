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
        self.loss = L.Loss(c.model.loss_list, c.model.vis_loss_list, num_gpu=c.num_gpu,
                           emulate_dist=bool(c.get("emulate_dist", False)))
        for n, p in self.model.named_parameters():
            if n.startswith(STATIC_UNUSED):
                p.requires_grad = False

    def forward(self, data_dict):
        data_dict = self.model(data_dict)
        total, all_losses = self.loss(data_dict)
        return total, all_losses


class PretrainStep:
    def __init__(self, cfg, device, total_steps=100000, dtype=torch.bfloat16, ddp=None, seed=0, cuda_graph=False,
                 overlap_allreduce=False):
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
        # several ranks, opt-in: the gradient all-reduce of everything BEHIND the text encoder (spatial layers, joint layers,
        # heads, ~57 % of the 491 MB) is launched from an autograd hook as soon as those gradients are final and runs on NCCL's
        # stream under the text encoder's backward; both collectives are captured in the step's CUDA graph.  Parity with the
        # plain path is checked at 2 ranks (scripts/dp_overlap_2rank.py); the long-running bench hung with it on the 2-GPU box
        # (round 2), so the default is the un-overlapped graph | all-reduce | graph path.
        self.overlap_allreduce = bool(overlap_allreduce)
        self.overlapped = False
        self._ar_work = None
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
        if self.flat_grads is not None:
            self.flat_grads.zero()
        else:
            self.optimizer.zero_grad(set_to_none=False)
        if self.flat_grads is None or not isinstance(self.flat_grads, FlatState):
            ops.refresh_shadows()
        # gradients live in the flat buffer: the wgrad kernels accumulate into it directly (no AccumulateGrad adds)
        prev, ops.DIRECT_GRAD[0] = ops.DIRECT_GRAD[0], self.flat_grads is not None
        self._ar_work = None
        self.module.model._lang_grad_hook = self._early_allreduce if self._overlap_now else None
        try:
            with torch.autocast(self.device.type, dtype=self.dtype, enabled=self.dtype != torch.float32):
                total, _ = self.module(dict(self.static_batch))
            total.backward()
        finally:
            ops.DIRECT_GRAD[0] = prev
            self.module.model._lang_grad_hook = None
        return total.detach()

    _overlap_now = False

    def _early_allreduce(self, grad):
        """autograd hook on the text encoder's output: its gradient is complete, so (the autograd engine runs nodes in reverse
        forward order) every node created after the text encoder's forward has run its backward — the gradients of the
        spatial layers, joint layers, heads and losses are final.  Their slice of the flat buffer goes out now."""
        if self._ar_work is None:
            tail = self.flat_grads.flat[self._lang_end:]
            self._ar_work = dist.all_reduce(tail, op=dist.ReduceOp.AVG, async_op=True)
        return None

    def _finish_allreduce(self):
        head = self.flat_grads.flat[:self._lang_end]
        if self._ar_work is not None:
            self._ar_work.wait()
            dist.all_reduce(head, op=dist.ReduceOp.AVG)
        else:                                   # the hook did not fire (no gradient reached the text encoder's output)
            self.flat_grads.all_reduce_mean()
        self._ar_work = None

    def _raw_opt(self):
        if isinstance(self.flat_grads, FlatState):
            self.flat_grads.optimizer_step()                    # native: clip + AdamW + bf16 shadows, one pass
            return
        if self.grad_norm is not None:
            if self.flat_grads is not None:
                self.flat_grads.clip_norm_(self.grad_norm)     # same math as clip_grad_norm_, two kernels on the flat buffer
            else:
                torch.nn.utils.clip_grad_norm_(self.parameters(), self.grad_norm, foreach=True)
        self.optimizer.step()

    def _raw_step(self):
        loss = self._raw_fwd_bwd()
        if self.flat_grads is not None:
            if self._overlap_now:
                self._finish_allreduce()
            else:
                self.flat_grads.all_reduce_mean()
        self._raw_opt()
        return loss

    def _lang_boundary(self):
        """Offset in the flat buffers where the text encoder's parameters end — valid only if they form the head of the layout
        (they do: model.get_opt_params lists the language encoder first)."""
        flat = self.flat_grads
        lang = {id(p) for p in self.module.model.lang_encoder.parameters()}
        end_lang = max((flat.offsets[i] + p.numel() for p in flat.params for i in [id(p)] if i in lang), default=0)
        first_other = min((flat.offsets[id(p)] for p in flat.params if id(p) not in lang), default=flat.n)
        if end_lang == 0 or end_lang > first_other:
            return None
        return (end_lang + 7) // 8 * 8 if (end_lang + 7) // 8 * 8 <= first_other else end_lang

    def _advance_lr(self):
        self._sched_step += 1
        f = self._lr_lambda(self._sched_step)
        if isinstance(self.flat_grads, FlatState):
            self.flat_grads.lr_factor.fill_(f)
            return
        for g, base in zip(self.optimizer.param_groups, self._base_lrs):
            g['lr'].fill_(base * f)

    def _capture(self, data_dict):
        self.static_batch = {k: v.clone() for k, v in data_dict.items() if torch.is_tensor(v)}
        if self.dp_graph:
            sync_module_state(self.module)
        ops.clear_shadows()
        kw = dict(self.cfg.solver.optim.args)
        self.flat_grads = FlatState(self.optimizer.param_groups, self._base_lrs, betas=tuple(kw.get("betas", (0.9, 0.999))),
                                    eps=float(kw.get("eps", 1e-8)), max_norm=self.grad_norm)
        self.flat_grads.lr_factor.fill_(self._lr_lambda(self._sched_step))
        register_packs(self.module, self.flat_grads)
        if self.dtype == torch.bfloat16:
            ops.register_shadows(self.module)       # whatever bf16 linears read outside the flat buffer (frozen parameters)
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
        if self.dp_graph and self.overlap_allreduce:
            self._lang_end = self._lang_boundary()
            if self._lang_end is not None:
                try:
                    self._overlap_now = True
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):       # one eager step through the hook path (NCCL warm-up on its stream)
                        self._raw_step()
                        self._advance_lr()
                    cur.wait_stream(side)
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=self.graph.pool()):
                        loss = self._raw_step()       # forward + backward + both all-reduces + clip/AdamW: ONE graph
                    self.graph_overlap, self.static_loss_overlap = g, loss
                    self.overlapped = True
                except Exception as e:                  # NCCL capture unsupported here: keep the two-graph path
                    self._overlap_now = False
                    torch.cuda.synchronize(self.device)
                    if dist.get_rank() == 0:
                        print(f"[sceneverse_b200] overlapped all-reduce unavailable ({type(e).__name__}: {e}); "
                              "using graph | all-reduce | graph")
                finally:
                    self._overlap_now = False

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
        if self.overlapped:
            self.graph_overlap.replay()
            self._advance_lr()
            return self.static_loss_overlap
        self.graph.replay()
        if self.dp_graph:
            self.flat_grads.all_reduce_mean()
            self.graph_opt.replay()
        self._advance_lr()
        return self.static_loss

    def phase_times(self, data_dict, steps=5):
        """(ms_fwd_bwd, ms_allreduce, ms_opt) of the NON-overlapped data-parallel step (graph | NCCL | graph), CUDA events on
        the current stream; the parameters do advance (these are real steps)."""
        assert self.dp_graph and self.graph is not None and self.graph_opt is not None
        warm = 3                      # un-timed: the ranks fall into step (the exchange kernel waits for the slowest peer)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(warm + steps)]
        for k, dst in self.static_batch.items():
            if data_dict[k].data_ptr() != dst.data_ptr():
                dst.copy_(data_dict[k], non_blocking=True)
        torch.cuda.synchronize(self.device)
        dist.barrier()
        for e in ev:
            e[0].record()
            self.graph.replay()
            e[1].record()
            self.flat_grads.all_reduce_mean()
            e[2].record()
            self.graph_opt.replay()
            e[3].record()
            self._advance_lr()
        torch.cuda.synchronize(self.device)
        return tuple(sum(e[i].elapsed_time(e[i + 1]) for e in ev[warm:]) / steps for i in range(3))

    def step(self, data_dict):
        """data_dict: tensors already on self.device. Returns the (detached) total loss tensor — no host sync.
        Graph mode: (1) the FIRST call runs the un-timed set-up on its batch — one probing forward/backward, three eager
        optimisation steps (they allocate gradient / optimizer state outside the graph pool), the capture, and the first replay —
        i.e. four parameter updates on that batch, unlike the reference's one step per batch; feed a throw-away batch first if
        that matters.  (2) the returned tensor is the graph's STATIC loss buffer, overwritten by the next replay: read it
        (`.item()` / `.clone()`) before calling step() again if the value has to be kept."""
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


class FlatState(FlatGrads):
    """Parameters, gradients, AdamW moments and the bf16 weight shadows of the trainable parameters as slices of FIVE flat
    buffers with one layout (parameter-group by parameter-group, every tensor starting on a 16-byte boundary of the bf16
    copy).  What that buys:
      * clip_grad_norm_ + AdamW + the fp32 -> bf16 shadow refresh are ONE pass over HBM (csrc/train_ops.cu adamw_flat_kernel:
        read p, g, m, v; write p, m, v, bf16 p) after a squared-norm pass over g — instead of torch's multi-tensor AdamW
        (16 launches), 2 clipping launches and a 65-launch shadow copy;
      * the data-parallel mean is one NCCL call on `flat` (as before);
      * sibling projections that read the same input (w_qs | w_ks | w_vs of the spatial attention, query | key | value of
        BERT) are ADJACENT in every buffer, so [3E, E] views of the shadow / gradient buffers let ops.linear_packed run them
        as one GEMM in each direction."""
    ALIGN = 8

    def __init__(self, groups, base_lrs, betas=(0.9, 0.999), eps=1e-8, max_norm=None):
        params, segs, off = [], [], 0
        for g, lr in zip(groups, base_lrs):
            ps = [p for p in g['params'] if p.requires_grad]
            if not ps:
                continue
            begin = off
            for p in ps:
                params.append((p, off))
                off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            segs.append((begin, off, float(lr), float(g.get('weight_decay', 0.0))))
        dev = params[0][0].device
        self.params = [p for p, _ in params]
        self.n = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)          # gradients (name kept from FlatGrads)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(off, dtype=torch.bfloat16, device=dev)
        self.offsets = {}
        pairs = []
        for p, o in params:
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[o:o + n].view_as(p)
            p.grad = self.flat[o:o + n].view_as(p)
            pairs.append((p, self.shadow[o:o + n].view_as(p)))
            self.offsets[id(p)] = o
        self.shadow.copy_(self.flat_p)
        ops.register_shadow_views(pairs)
        import numpy as np
        seg_dt = np.dtype([("begin", "<i8"), ("end", "<i8"), ("lr_scale", "<f4"), ("weight_decay", "<f4")])
        self.segments = torch.from_numpy(np.array(segs, dtype=seg_dt).view(np.uint8).copy()).to(dev)
        self.nseg = len(segs)
        self.betas, self.eps, self.max_norm = betas, eps, max_norm
        self.lr_factor = torch.ones(1, dtype=torch.float32, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.grad_norm_t = torch.zeros(1, dtype=torch.float32, device=dev)
        from . import _lib
        self.scratch = torch.zeros(_lib.gps().sv_adamw_scratch_floats(), dtype=torch.float32, device=dev)

    def packed_views(self, params):
        """[sum N, K] / [sum N] views of the shadow, gradient and parameter buffers spanning `params` if they are adjacent
        in the flat layout (same trailing shape, no alignment gaps), else None."""
        offs = [self.offsets.get(id(p)) for p in params]
        if any(o is None for o in offs):
            return None
        tail = params[0].shape[1:]
        o = offs[0]
        for p, po in zip(params, offs):
            if po != o or p.shape[1:] != tail:
                return None
            o += p.numel()
        rows = sum(p.shape[0] for p in params)
        sl = slice(offs[0], o)
        return (self.shadow[sl].view(rows, *tail), self.flat[sl].view(rows, *tail), self.flat_p[sl].view(rows, *tail))

    def optimizer_step(self):
        """clip (max_norm) + AdamW + shadow refresh; the step counter lives on the device (CUDA-graph replay safe)."""
        from . import _lib
        lib = _lib.gps()
        self.step_t.add_(1)
        with torch.cuda.device(self.flat.device):
            st = lib.sv_adamw_flat(self.flat_p.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                   self.flat.data_ptr(), self.shadow.data_ptr(), self.n, self.segments.data_ptr(), self.nseg,
                                   float(self.max_norm) if self.max_norm is not None else 0.0, self.lr_factor.data_ptr(),
                                   self.step_t.data_ptr(), 1.0, float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                   self.scratch.data_ptr(), self.grad_norm_t.data_ptr(),
                                   torch.cuda.current_stream(self.flat.device).cuda_stream)
        _lib.check(lib, st, "sv_adamw_flat")


def register_packs(module, flat):
    """Tell ops.linear_packed which sibling projections are adjacent in the flat buffers."""
    from .modules.layers import MultiHeadAttentionSpatial
    for m in module.modules():
        trip = None
        if isinstance(m, MultiHeadAttentionSpatial):
            trip = (m.w_qs, m.w_ks, m.w_vs)
        elif all(hasattr(m, a) and isinstance(getattr(m, a), nn.Linear) for a in ("query", "key", "value")):
            trip = (m.query, m.key, m.value)       # HF BertSelfAttention
        if trip is None or any(l.bias is None for l in trip):
            continue
        w = flat.packed_views([l.weight for l in trip])
        b = flat.packed_views([l.bias for l in trip])
        if w is not None and b is not None:
            ops.register_pack([l.weight for l in trip], [l.bias for l in trip], w, b)


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


class ObjClsStep:
    """Object-level pre-training step (reference: model/objcls.py:64-98 + optim/loss/loss.py:96-102, BASELINE.json configs[1]):
    PointNet++ TRAINABLE with train-mode BatchNorm -> dropout -> 607-way open-vocabulary logits -> label-smoothed CE ->
    AdamW.  The point operators and their gradients (group_points_grad / gather) are the native kernels; conv / BatchNorm
    statistics run on cuDNN / ATen (the fused tensor-core SA-MLP is the eval-BN form).  cuda_graph=True captures the whole
    step after three eager ones."""

    def __init__(self, device, text_embeds, lr=1e-3, dtype=torch.bfloat16, seed=0, cuda_graph=False, num_gpu=1):
        torch.manual_seed(seed)
        self.device, self.dtype = torch.device(device), dtype
        self.model = M.ObjCls({"num_gpu": num_gpu, "solver": {"lr": lr}}, text_embeds=text_embeds).to(self.device).train()
        self.graph_mode = bool(cuda_graph) and self.device.type == "cuda"
        self.optimizer = torch.optim.AdamW(self.model.parameters(), lr=torch.tensor(lr, device=self.device) if self.graph_mode else lr,
                                           fused=self.device.type == "cuda", capturable=self.graph_mode)
        self.graph = self.static_batch = self.static_loss = None

    def _raw(self, batch):
        self.optimizer.zero_grad(set_to_none=False)
        with torch.autocast(self.device.type, dtype=self.dtype, enabled=self.dtype != torch.float32):
            out = self.model(dict(batch))
            loss = L.obj_cls_loss(out)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def step(self, batch):
        if not self.graph_mode:
            return self._raw(batch)
        if self.graph is None:
            self.static_batch = {k: v.clone() for k, v in batch.items() if torch.is_tensor(v)}
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._raw(self.static_batch)
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            from . import _lib
            n0 = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._raw(self.static_batch)
            self.native_launches_per_step = _lib.launch_count() - n0
        for k, dst in self.static_batch.items():
            if batch[k].data_ptr() != dst.data_ptr():
                dst.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.static_loss
