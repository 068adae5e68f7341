"""Fused normalise + all-gather over NVLink peer memory (csrc/norm_allgather.cu; reference: common/dist_utils.py:131-149
+ the F.normalize calls of optim/loss/contra_loss.py:58-64,86-91).
* single GPU: a world = 1 self-exchange through the same kernel, epoch flags and parity logic, eager and under CUDA-graph
  replay with an ODD number of exchanges per replay (the case the host-side parity of round 1 got wrong);
* >= 2 GPUs: two ranks against F.normalize + NCCL all_gather (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_self_exchange_matches_normalize_eager_and_graph():
    from sceneverse_b200 import fused_gather
    dev = torch.device("cuda", 0)
    fg = fused_gather.FusedNormGather(64, 768, dev, local=True)
    g = torch.Generator(device=dev).manual_seed(5)
    for it in range(5):                         # eager: parity alternates call by call
        a = torch.randn(64, 768, device=dev, generator=g) * (it + 1)
        b = torch.randn(64, 768, device=dev, generator=g)
        if it == 3:
            a[7].zero_()                        # F.normalize eps path: an all-zero row stays zero
        ga, gb = fg(a, b)
        assert torch.equal(ga, F.normalize(a, dim=-1)) or (ga - F.normalize(a, dim=-1)).abs().max() < 1e-6
        assert (gb - F.normalize(b, dim=-1)).abs().max() < 1e-6
        assert ga.data_ptr() != fg.buf.data_ptr() and not ga.requires_grad   # a copy, detached (reference semantics)
    assert int(fg.epoch_dev) == 5
    # graph replay, ONE exchange per replay: the device epoch (and so the parity) advances on every replay
    sa = torch.randn(64, 768, device=dev, generator=g)
    sb = torch.randn(64, 768, device=dev, generator=g)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fg(sa, sb)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        oa, ob = fg(sa, sb)
    e0 = int(fg.epoch_dev)
    for it in range(5):
        sa.copy_(torch.randn(64, 768, device=dev, generator=g) * 3)
        sb.copy_(torch.randn(64, 768, device=dev, generator=g))
        graph.replay()
        torch.cuda.synchronize()
        assert (oa - F.normalize(sa, dim=-1)).abs().max() < 1e-6, it
        assert (ob - F.normalize(sb, dim=-1)).abs().max() < 1e-6, it
    assert int(fg.epoch_dev) == e0 + 5
    # and an eager call after the replays still reads the half the kernel wrote
    ga, _ = fg(sa * 2, sb)
    assert (ga - F.normalize(sa, dim=-1)).abs().max() < 1e-6


def test_fused_normalize_allgather_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577",
                          os.path.join(ROOT, "scripts", "fused_allgather_2rank.py")], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert "FUSED_ALLGATHER_OK=True" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_overlapped_gradient_allreduce_two_ranks():
    """train.PretrainStep at 2 ranks: all-reduce split at the text encoder boundary, launched from an autograd hook and
    captured in the step graph, vs the un-overlapped step (needs 2 GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29578",
                          os.path.join(ROOT, "scripts", "dp_overlap_2rank.py")], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert "DP_OVERLAP_OK=True" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
