"""Multi-GPU host logic: one process per GPU, chains sharded by contiguous global id ranges
(SURVEY.md 8e).  The step loop has no collective; `torch.distributed` (NCCL over NVLink on the
GPU box, gloo in the CPU tests) carries only the small sum-reductions of the two optional
cross-chain features -- pooled adaptation and R-hat/ESS -- through the C ABI's allreduce
callback (include/mcmcb200.h, mcmcb_set_allreduce).

PyTorch is plumbing here: process group, streams, and a tensor view of the library's device
buffer.  Nothing in this file computes on the chains.
"""
import os


def env_rank():
    """(rank, world_size, local_rank) as torchrun exports them; (0, 1, 0) outside torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard(nchains_total, rank, world):
    """Contiguous chain-id range of `rank`: (nchains_local, chain_offset).  The Philox stream of a
    chain is keyed by its GLOBAL id, so results do not depend on `world` (pooling off)."""
    if not (0 <= rank < world) or nchains_total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(int(nchains_total), int(world))
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return n, off


class _DevicePtr:
    """__cuda_array_interface__ view of n doubles at a raw device pointer (no copy, no ownership)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def make_allreduce(group=None):
    """Callable (device_ptr, n, cuda_stream) -> None that sum-reduces n doubles in place across `group`.

    NCCL: the reduction is enqueued behind the library's kernels on the library's own stream (wrapped as a
    torch ExternalStream), so the host never blocks.  gloo (no GPU-side collective): the buffer is staged
    through the host after a stream synchronise."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None  # local sums are the global sums
    backend = dist.get_backend(group)

    def _allreduce(ptr, n, stream):
        dev = torch.cuda.current_device()
        t = torch.as_tensor(_DevicePtr(ptr, n), device=torch.device("cuda", dev))
        ext = torch.cuda.ExternalStream(stream, device=dev) if stream else torch.cuda.current_stream(dev)
        with torch.cuda.stream(ext):
            if backend == "nccl":
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                ext.synchronize()
                h = t.cpu()
                dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
                t.copy_(h)
                ext.synchronize()

    return _allreduce


def attach(sampler, group=None):
    """Give a Sampler the job-wide allreduce (pool_adapt / diagnostics become collective)."""
    fn = make_allreduce(group)
    sampler.set_allreduce(fn)
    return fn is not None


def allreduce_host(array, group=None):
    """Sum-reduce a host numpy array of doubles in place across `group` (what the gloo staging path of
    make_allreduce does with the library's buffer); returns the array."""
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.from_numpy(array)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return array
