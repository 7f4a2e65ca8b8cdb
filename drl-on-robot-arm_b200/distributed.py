"""Multi-GPU plumbing: one process per GPU (torchrun), envs sharded by contiguous blocks, and the ONLY exchange on
the training path -- one all-reduce of a flat gradient bucket per network per optimizer step.

The simulation kernel needs no collective: every env is independent and its Philox stream is keyed by the GLOBAL env
id (ArmsimConfig.env_id_offset), so results do not depend on how the batch is cut into ranks.
The gradient buckets are 0.27-0.55 MB (TD3 twin critic 137 218 params, actor 68 355): latency-bound messages, so each
network's gradients live in ONE contiguous buffer (param.grad are views into it) and go out as ONE all-reduce.
Works with backend "nccl" (NVLink/NVSwitch) and "gloo" (CPU tests).
"""
import gc
import os
import weakref

import torch
import torch.distributed as dist

# objects holding CUDA graphs that captured NCCL collectives (VectorTrainer registers itself): shutdown() makes them drop
# those graphs first -- ncclCommDestroy blocks for ever while an instantiated graph still references the communicator
# (measured on 2 x B200, torch 2.11 / NCCL 2.28: the run finished, then hung in destroy_process_group)
_GRAPH_HOLDERS = weakref.WeakSet()


def register_graph_holder(obj):
    """obj.release_graphs() will be called by shutdown() before the process group is destroyed"""
    _GRAPH_HOLDERS.add(obj)


def shutdown():
    """ordered teardown of a multi-process run: drop recorded collectives, drain the device, barrier, destroy the group"""
    for h in list(_GRAPH_HOLDERS):
        h.release_graphs()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  Returns (rank, world,
    local_rank).  Single-process runs (no WORLD_SIZE) return (0, 1, 0) without creating a group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_range(n_total, rank, world):
    """rank r owns envs [lo, hi): contiguous blocks, remainder spread over the first ranks."""
    if n_total < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad shard request n=%d rank=%d world=%d" % (n_total, rank, world))
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def make_sharded_env(task, n_total, rank=None, world=None, device=None, **kw):
    """This rank's shard of an n_total-env job as a BatchedArmEnv whose env ids are global."""
    from .envs import BatchedArmEnv
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(n_total, rank, world)
    kw.pop("env_id_offset", None)
    return BatchedArmEnv(task, n_envs=hi - lo, device=device, env_id_offset=lo, **kw)


class GradBucket:
    """All gradients of one network in one contiguous buffer; `p.grad` are views into it.

    usage:   bucket = GradBucket(net.parameters())
             bucket.zero(); loss.backward(); bucket.allreduce_mean(); optimizer.step()
    (use bucket.zero() instead of optimizer.zero_grad(), which would drop the views)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.group = group
        self._work = None

    def zero(self):
        self.flat.zero_()
        for p in self.params:                   # re-attach if something replaced .grad
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * self.flat.element_size():
                self._reattach()
                break

    def _reattach(self):
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def allreduce_mean(self, async_op=False):
        """sum over ranks / world, in place, ONE collective.  No-op for a single process."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return None
        world = dist.get_world_size(self.group)
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        else:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if async_op:
                work.wait()
                work = None
            self.flat.div_(world)
        self._work = work
        return work

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


def broadcast_module(module, src=0, group=None):
    """make every replica start from rank `src`'s weights (one flat broadcast)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    ps = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    flat = torch.cat([t.reshape(-1) for t in ps])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for t in ps:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()


def allreduce_scalars(values, op="sum", group=None):
    """tiny all-reduce for logging (success counts, returns): dict name -> float, one collective for all of them"""
    keys = sorted(values)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(values)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(values[k]) for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op], group=group)
    return {k: float(v) for k, v in zip(keys, t.tolist())}
