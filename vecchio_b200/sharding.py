"""Samples-per-pixel sharding of one frame over N ranks: the only multi-GPU logic the path
has.  Samples are i.i.d. and the pixel is a plain mean (src/main.rs:186-197), so rank k renders the
global sample indices [k*spp/N, (k+1)*spp/N) of EVERY pixel into a buffer of per-pixel SUMS, the
buffers are combined by one reduce(sum) to rank 0, and rank 0 divides by spp once (dropped samples
stay in the divisor).  The Philox counter holds the global sample index, so the image does not
depend on N up to the fp32 order of the N partial sums."""


def spp_slice(rank, world, spp):
    """(spp_begin, spp_count) of `rank`; the slices tile [0, spp) exactly, sizes differ by at most 1."""
    if not (0 <= rank < world) or spp < world:
        raise ValueError("need 0 <= rank < world <= spp")
    begin, end = rank * spp // world, (rank + 1) * spp // world
    return begin, end - begin


def reduce_sums_to_root(sum_tensor, world):
    """One reduce(sum) of the fp32 per-pixel sums to rank 0 (NCCL over NVLink on the GPUs, gloo in the
    CPU tests), enqueued on the current stream; a no-op for a single rank."""
    if world > 1:
        import torch.distributed as dist
        dist.reduce(sum_tensor, dst=0, op=dist.ReduceOp.SUM)
    return sum_tensor
