"""normalise + all-gather over NVLink peer memory (csrc/norm_allgather.cu) on top of torch's symmetric-memory
allocator (the rendezvous / peer mapping is torch.distributed plumbing; the exchange itself is one native kernel)."""
import torch
import torch.distributed as dist

from . import _lib

_state = {}


class FusedNormGather:
    SIG_WORDS = 64

    def __init__(self, n, D, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n, self.D = n, D
        self.buf = symm_mem.empty(2 * 2 * self.world * n * D + self.SIG_WORDS, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        data_floats = 2 * 2 * self.world * n * D
        self.peer_bufs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.peer_sigs = torch.tensor([p + 4 * data_floats for p in ptrs], dtype=torch.int64, device=device)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=device)  # the kernel's own epoch word (graph-replay safe)
        self.epoch = 0                                                       # host mirror: selects the parity of the views
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every rank's buffer (incl. signal words) is zeroed before the first exchange

    def __call__(self, a, b):
        """a, b: (n,D) -> (normalize(a) gathered (world*n,D), normalize(b) gathered), views of the symmetric buffer.
        The views alternate between the two parity halves of the buffer call by call; inside a replayed CUDA graph the
        parity of a call site is the one it had at capture, which matches the device epoch as long as the number of
        exchanges per replay is even (two between-batch losses in the pre-training configuration)."""
        a = a.detach().float().contiguous()
        b = b.detach().float().contiguous()
        assert a.shape == (self.n, self.D) and b.shape == (self.n, self.D)
        self.epoch += 1
        lib = _lib.gps()
        with torch.cuda.device(a.device):
            st = lib.sv_normalize_allgather_dev_f32(a.data_ptr(), b.data_ptr(), self.n, self.D, self.peer_bufs.data_ptr(),
                                                    self.peer_sigs.data_ptr(), self.done.data_ptr(), self.world, self.rank,
                                                    self.epoch_dev.data_ptr(), torch.cuda.current_stream(a.device).cuda_stream)
        _lib.check(lib, st, "sv_normalize_allgather_dev_f32")
        wn = self.world * self.n
        base = (self.epoch & 1) * 2 * wn * self.D
        ga = self.buf[base: base + wn * self.D].view(wn, self.D)
        gb = self.buf[base + wn * self.D: base + 2 * wn * self.D].view(wn, self.D)
        return ga, gb


def get(n, D, device):
    """One exchange object per (n, D, device); None if symmetric memory cannot be set up on this system."""
    key = (n, D, str(device))
    if key not in _state:
        try:
            _state[key] = FusedNormGather(n, D, device)
        except Exception as e:  # no P2P / unsupported allocator: the NCCL path of losses.all_gather is used instead
            _state[key] = None
            if dist.get_rank() == 0:
                print(f"[sceneverse_b200] fused normalise+all-gather unavailable ({type(e).__name__}: {e}); using NCCL all_gather")
    return _state[key]
