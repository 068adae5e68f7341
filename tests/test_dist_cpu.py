"""World-size-2 gloo tests (CPU) of the two collectives of the GPS path: the embedding all-gather inside the
batch-contrastive losses (reference semantics: rank-major concat, gathered tensors carry no gradient) and the
data-parallel gradient mean."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sceneverse_b200.modules import losses
        g = torch.Generator().manual_seed(100 + rank)
        B, O, D = 4, 6, 32
        obj = torch.randn(B, O, D, generator=g, requires_grad=True)
        txt = torch.randn(B, D, generator=g, requires_grad=True)
        tgt = torch.randint(0, O, (B, 1), generator=g)
        # 1. all_gather: rank-major concatenation, detached
        a, b = losses.all_gather([obj[:, 0], txt])
        assert a.shape == (world * B, D) and not a.requires_grad and not b.requires_grad
        assert torch.equal(a[rank * B:(rank + 1) * B], obj[:, 0].detach())
        # 2. distributed InfoNCE == single-process InfoNCE on the concatenated batch
        loss_d = losses.TextObjBetweenBatch({"num_gpu": world})
        val = loss_d({"inter_obj_embeds": obj, "inter_text_embed": txt, "tgt_object_id": tgt})
        val.backward()
        assert obj.grad is None or float(obj.grad.abs().sum()) == 0.0  # reference quirk: no grad through the gather
        assert loss_d.logit_scale.grad is not None
        objs, txts, tgts = [], [], []
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            objs.append(torch.randn(B, O, D, generator=gr))
            txts.append(torch.randn(B, D, generator=gr))
            tgts.append(torch.randint(0, O, (B, 1), generator=gr))
        single = losses.TextObjBetweenBatch({"num_gpu": 1})
        want = single({"inter_obj_embeds": torch.cat(objs), "inter_text_embed": torch.cat(txts),
                       "tgt_object_id": torch.cat(tgts)})
        assert abs(float(val) - float(want)) < 1e-5, (float(val), float(want))
        # 3. DDP gradient mean of a replicated module
        lin = torch.nn.Linear(D, 3)
        torch.manual_seed(0)
        for p in lin.parameters():
            torch.nn.init.normal_(p)
        ddp = torch.nn.parallel.DistributedDataParallel(lin)
        x = torch.randn(5, D, generator=g)
        ddp(x).pow(2).sum().backward()
        gathered = [torch.zeros_like(lin.weight.grad) for _ in range(world)]
        dist.all_gather(gathered, lin.weight.grad)
        assert torch.allclose(gathered[0], gathered[1])
        # 4. the graph-mode data-parallel path: flat gradient buffer + one all-reduce (mean) == DDP's gradient mean
        from sceneverse_b200 import train
        lin2 = torch.nn.Linear(D, 3)
        if rank == 1:
            with torch.no_grad():
                for p in lin2.parameters():
                    p.add_(1.0)                       # deliberately out of sync: sync_module_state must repair it
        train.sync_module_state(lin2)
        with torch.no_grad():
            for p2, p1 in zip(lin2.parameters(), lin.parameters()):
                p2.copy_(p1)
        fg = train.FlatGrads(list(lin2.parameters()))
        for _ in range(2):                            # the second pass checks zero() + in-place accumulation into the views
            fg.zero()
            lin2(x).pow(2).sum().backward()
            assert lin2.weight.grad.data_ptr() == fg.flat.data_ptr()
            fg.all_reduce_mean()
            assert torch.allclose(lin2.weight.grad, lin.weight.grad, atol=1e-6), (lin2.weight.grad, lin.weight.grad)
            assert torch.allclose(lin2.bias.grad, lin.bias.grad, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
