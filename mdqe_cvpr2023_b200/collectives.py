"""Gradient all-reduce over NVLink peer memory: the host side of csrc/allreduce.cu.

The reference trains with PyTorch DDP over NCCL (train_net.py:256-271).  Here the bucket lives in a *symmetric* allocation --
every rank maps every other rank's copy, plus the NVSwitch multicast alias when the box has one -- and ONE small kernel of this
library reduces it in place (`msda_allreduce_f32`): multimem.ld_reduce / multimem.st through the switch, or two-shot P2P loads
and stores.  torch.distributed is used for the plumbing only: allocation and handle exchange
(`torch.distributed._symmetric_memory`), never for the reduction itself.

    ar = PeerAllReduce(n_floats, device)            # collective: every rank of the group calls it
    ar.buffer[...] = flat gradients                 # (or build the gradient buckets as views of ar.buffer)
    ar.all_reduce_(mean=True)                       # enqueue on the current stream; graph-capturable
    ar.check()                                      # after a synchronize: raises if a peer never arrived
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class PeerAllReduce:
    ALGOS = {"p2p": 0, "multimem": 1}

    def __init__(self, numel, device, group=None, algo="auto", n_ctas=8):
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized():
            raise RuntimeError("PeerAllReduce needs an initialised process group (plumbing: handle exchange)")
        lib = _lib.load()
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > lib.msda_allreduce_max_ranks():
            raise RuntimeError(f"PeerAllReduce supports up to {lib.msda_allreduce_max_ranks()} ranks (one NVSwitch box)")
        self.numel = (int(numel) + 3) // 4 * 4
        self.n_ctas = int(n_ctas)
        self.device = torch.device(device)
        self.buffer = symm.empty(self.numel, dtype=torch.float32, device=self.device)
        self._flags = symm.empty(lib.msda_allreduce_flag_bytes(148) // 4, dtype=torch.int32, device=self.device)
        self.buffer.zero_()
        self._flags.zero_()
        self._hb = symm.rendezvous(self.buffer, group.group_name)
        self._hf = symm.rendezvous(self._flags, group.group_name)
        self._peers = (ctypes.c_uint64 * self.world)(*[int(p) for p in self._hb.buffer_ptrs])
        self._flag_ptrs = (ctypes.c_uint64 * self.world)(*[int(p) for p in self._hf.buffer_ptrs])
        self._mc = int(getattr(self._hb, "multicast_ptr", 0) or 0)          # 0: no NVSwitch multicast on this box
        if algo == "auto":
            algo = "multimem" if self._mc else "p2p"
        if algo == "multimem" and not self._mc:
            raise RuntimeError("multimem all-reduce needs NVSwitch multicast support on this box")
        self.algo = algo
        self._error = torch.zeros(1, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(group)                          # every rank's flags are zero before anybody's first kernel

    def all_reduce_(self, offset=0, numel=None, mean=True, stream=None, n_ctas=None):
        """bucket[offset:offset+numel] = (mean or sum) over the ranks, in place in every replica.  Enqueues one kernel."""
        numel = self.numel - offset if numel is None else int(numel)
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        _lib.check(_lib.load().msda_allreduce_f32(st, self.ALGOS[self.algo], self.rank, self.world, self._peers, self._mc, self._flag_ptrs,
                                                  self._error.data_ptr(), int(offset), numel, (1.0 / self.world) if mean else 1.0,
                                                  n_ctas or self.n_ctas), "msda_allreduce_f32")
        return self.buffer[offset:offset + numel]

    def check(self):
        """after a device synchronize: did every peer arrive at every barrier?"""
        if int(self._error.item()) != 0:
            raise RuntimeError("msda_allreduce_f32: a peer never reached the barrier (bounded spin expired)")


def allreduce_mean_gradients_peer(parameters, ar):
    """DDP's gradient averaging with the peer-memory kernel: every parameter's .grad (zeros where this rank produced none) is
    gathered into ar.buffer with ONE concatenation kernel, reduced in place by msda_allreduce_f32, and scattered back with one
    multi-tensor copy.  Same collective on every rank, like sharding.allreduce_mean_gradients; graph-capturable."""
    params = [p for p in parameters if p.requires_grad]
    for p in params:
        if p.dtype != torch.float32:
            raise RuntimeError("allreduce_mean_gradients_peer: fp32 parameters only (the bucket is an fp32 symmetric buffer)")
    total = sum(p.numel() for p in params)
    if total > ar.numel:
        raise RuntimeError(f"bucket of {ar.numel} floats is too small for {total} gradient elements")
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    n4 = (total + 3) // 4 * 4
    flat = ar.buffer[:total]
    torch.cat([p.grad.reshape(-1) for p in params], out=flat)
    if n4 > total:
        ar.buffer[total:n4].zero_()
    ar.all_reduce_(0, n4, mean=True)
    pieces = flat.split([p.numel() for p in params])
    torch._foreach_copy_([p.grad.view(-1) for p in params], list(pieces))
    return 1


def make_ddp_comm_hook(ar):
    """A `torch.nn.parallel.DistributedDataParallel` communication hook that averages every gradient bucket with the peer-memory
    kernel instead of NCCL's all-reduce -- how the reference's training loop (detectron2 wraps the model in DDP, train_net.py:256-271)
    picks the kernel up without any other change:

        ar = PeerAllReduce(largest_bucket_elements, device)
        ddp_model.register_comm_hook(None, make_ddp_comm_hook(ar))

    The bucket is copied into the symmetric buffer, reduced in place and copied back, all on the current stream; the returned future
    is already complete (its tensor carries the stream-ordered result, like DDP's own no-op hook pattern)."""
    def hook(state, bucket):
        t = bucket.buffer()
        n = t.numel()
        if t.dtype != torch.float32 or n > ar.numel:
            raise RuntimeError(f"peer all-reduce hook: bucket of {n} {t.dtype} elements does not fit the fp32 buffer of {ar.numel}")
        n4 = (n + 3) // 4 * 4
        flat = t.reshape(-1)
        ar.buffer[:n].copy_(flat)
        if n4 > n:
            ar.buffer[n:n4].zero_()
        ar.all_reduce_(0, n4, mean=True)
        flat.copy_(ar.buffer[:n])
        fut = torch.futures.Future()
        fut.set_result(t)
        return fut
    return hook
