"""Parallel backend plumbing (host side): chain sharding and communicator set-up.

ref: src/Parallel/base_parallel.jl:1-9 (automatic_parallel_type), src/Parallel/MPI/mpi.jl:21-74,
     src/Samplers/Metropolis.jl:75-76 (per-worker chain length and seed).
The reference shards the chain LENGTH over MPI ranks and gives each rank its own seed; here ranks own
disjoint sets of CHAINS (the burn-in is not replicated) and every chain's Philox stream is keyed by its
global id, so results do not depend on the number of GPUs.  Collectives are NCCL all-reduces issued by
libnqcuda on the context's stream (nq_comm.cu); torch.distributed is only used to hand the NCCL unique
id to the other ranks.
"""
import os


def shard_chains(total_chains, nranks, rank):
    """(offset, count) of the chains owned by `rank`: contiguous blocks, remainder to the first ranks."""
    if not (0 <= rank < nranks):
        raise ValueError("rank %d outside world of %d" % (rank, nranks))
    base, rem = divmod(int(total_chains), int(nranks))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def world_from_env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_comm(ctx, dist=None):
    """Create the NCCL communicator of `ctx` for the torch.distributed world (no-op for one rank)."""
    from .core import unique_id
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 1, 0
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(world, rank, box[0])
    return world, rank
