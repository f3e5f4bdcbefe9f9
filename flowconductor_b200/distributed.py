"""Batch sharding across the GPUs of one box: rows are independent, every rank holds a full
model replica and a contiguous block of rows, and there is NO collective on the data path.  NCCL (over
NVLink / NVSwitch) is used only for
  * the final log-likelihood reduction (one fp64 scalar per call), and
  * the training-step gradient all-reduce (one flat bucket of all parameter gradients).
The reference has no distributed code; this is new host logic.  It works with any torch.distributed backend
(`nccl` on the GPU box, `gloo` in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(num_rows, rank, world_size):
    """Contiguous, balanced row block [lo, hi) owned by `rank` (first `num_rows % world_size` ranks get one
    extra row)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank {} outside world of size {}".format(rank, world_size))
    base, extra = divmod(num_rows, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(tensor, rank=None, world_size=None):
    """The rows of `tensor` owned by this rank."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    lo, hi = shard_bounds(tensor.shape[0], rank, world_size)
    return tensor[lo:hi]


def reduce_log_likelihood(local_log_prob, group=None):
    """(sum over ALL ranks of log_prob, total row count), both 0-dim fp64 DEVICE tensors: local fp64 sum, then one
    2-element all-reduce.  No host synchronisation: the caller decides when (and whether) to read them back
    (`float(total)`, `int(count)`)."""
    packed = torch.stack((local_log_prob.double().sum(),
                          torch.tensor(float(local_log_prob.numel()), dtype=torch.float64,
                                       device=local_log_prob.device)))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[0], packed[1]


def sharded_log_prob(flow, inputs, context=None, chunk_rows=None):
    """`flow.log_prob` over this rank's rows, optionally streamed in chunks (cfg 5: 100 M rows do not fit as
    one batch; the reference only offers chunking for sampling, distributions/base.py:69-81)."""
    if chunk_rows is None or inputs.shape[0] <= chunk_rows:
        return flow.log_prob(inputs, context=context)
    out = []
    for lo in range(0, inputs.shape[0], chunk_rows):
        ctx = None if context is None else context[lo:lo + chunk_rows]
        out.append(flow.log_prob(inputs[lo:lo + chunk_rows], context=ctx))
    return torch.cat(out)


def allreduce_gradients(module, group=None, average=True):
    """Flat-bucket gradient all-reduce for data-parallel training (cfg 3: 2.3 M parameters = 9.2 MB, one
    NCCL call per step).  Returns the number of elements reduced."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return 0
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset:offset + n].view_as(g))
        offset += n
    return offset


def broadcast_parameters(module, src=0, group=None):
    """Make every replica identical to rank `src` (parameters and buffers)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    from .nn import tensorcore

    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
    # the collective wrote through .data: neither the pointers nor the version counters of the parameters moved, so the
    # packed tensor-core weights cached on the conditioners would be stale
    tensorcore.invalidate(module)


def host_log_prob(flow, inputs_host, out_host=None, chunk_rows=262144, context_host=None):
    """`flow.log_prob` for a batch that lives in (pinned) HOST memory: rows are streamed to the GPU in chunks on a copy
    stream, double-buffered, so that the host-to-device copy of chunk i+1 overlaps the kernels of chunk i, and each
    chunk's log-probabilities are copied back as soon as they exist.  Returns the host tensor of log-probabilities
    (valid after the current stream has been synchronised).  No gradient."""
    dev = next(flow.parameters()).device
    n, feat = inputs_host.shape
    if out_host is None:
        out_host = torch.empty((n,), dtype=torch.float32).pin_memory()
    main = torch.cuda.current_stream(dev)
    copier = torch.cuda.Stream(dev)
    rows = min(chunk_rows, max(n, 1))
    xbuf = [torch.empty((rows, feat), dtype=torch.float32, device=dev) for _ in range(2)]
    cbuf = None
    if context_host is not None:
        cbuf = [torch.empty((rows, context_host.shape[1]), dtype=torch.float32, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    for ev in free:
        ev.record(main)
    # The first chunk is small — two waves of the persistent kernels (256 rows per CTA pair) — so that the pipeline fills in
    # ~0.2 ms instead of the copy time of a whole chunk; from then on the copier stays ahead of the kernels.
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    lead = min(rows, 2 * 256 * max(sms // 2, 1))
    bounds, lo = [], 0
    while lo < n:
        hi = min(n, lo + (lead if not bounds else rows))
        bounds.append((lo, hi))
        lo = hi
    with torch.no_grad():
        for i, (lo, hi) in enumerate(bounds):
            b = i & 1
            with torch.cuda.stream(copier):
                copier.wait_event(free[b])            # the kernels of chunk i-2 have finished with this buffer
                xbuf[b][:hi - lo].copy_(inputs_host[lo:hi], non_blocking=True)
                if cbuf is not None:
                    cbuf[b][:hi - lo].copy_(context_host[lo:hi], non_blocking=True)
                ready[b].record(copier)
            main.wait_event(ready[b])
            lp = flow.log_prob(xbuf[b][:hi - lo], context=None if cbuf is None else cbuf[b][:hi - lo])
            out_host[lo:hi].copy_(lp, non_blocking=True)
            free[b].record(main)
    return out_host
