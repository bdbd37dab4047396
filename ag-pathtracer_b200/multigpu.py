"""Sample-index sharding across the GPUs of one node.

Every (pixel, sample) path is independent given its RNG stream and accumulation is a sum, so
GPU g of G renders the samples s = g (mod G) of every pixel into its own float4[W*H]
accumulator (scene replicated, no ray migration, no data-path collective); ONE all-reduce of
the accumulators at the end produces the final framebuffer.  The only thing that differs
from a single-GPU render is the fp32 summation order.
"""


def shard(first_sample, num_samples, rank, world):
    """Samples first_sample + k, k in [0, num_samples), k = rank (mod world).
    Returns (first, count, stride) as agpt_render takes them."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    count = (num_samples - rank + world - 1) // world if num_samples > rank else 0
    return first_sample + rank, count, world


def render_sharded(ctx, first_sample, num_samples, max_depth, depth_arg, rank, world, flags=0):
    """This rank's share of the samples into ctx's accumulator (no communication)."""
    first, count, stride = shard(first_sample, num_samples, rank, world)
    if count > 0:
        ctx.render(first, count, max_depth, depth_arg, flags, sample_stride=stride)
    return count


def allreduce_accumulator(accum, group=None):
    """Sum the per-rank float4 accumulators (torch tensor) in place over NCCL / gloo."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
    return accum
