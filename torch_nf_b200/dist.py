"""Data-parallel sharding of the sample batch (one process per GPU).

The hot path shards by independent sample rows: every rank holds the full
weights and its slice of ``N`` (shared weights) or of ``M`` (per-sample
weights).  ``log_prob`` needs no communication.  The only exchange steps are
  * BatchNorm batch statistics in the sample direction
    (reference torch_nf/bijectors.py:401-415 pools over the WHOLE batch):
    one all-reduce of ``[sum(D) | sumsq(D) | count]`` float64 per BatchNorm,
    and of ``[sum g | sum g*y]`` in its backward;
  * the gradient all-reduce and the scalar loss / log-prob means in training.
Collectives go through ``torch.distributed`` (NCCL over NVLink on the GPU box,
gloo in the CPU tests).
"""
import torch

_group = None
_enabled = False


def enable(group=None):
    """Turn on cross-rank BatchNorm statistics / reductions for this process."""
    global _group, _enabled
    import torch.distributed as td
    if not td.is_available() or not td.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _group, _enabled = group, True


def disable():
    global _group, _enabled, _peer
    _group, _enabled, _peer = None, False, None


# ---- BatchNorm statistics over NVLink peer memory (tnf_peer_t, include/tnf.h) --------------------------------------
_peer = None          # dict(stats, flags, stats_ptrs, flag_ptrs, seq) once enable_peer_exchange() succeeded
peer_error = None     # why it did not (the NCCL all-reduce hook is used then)


def enable_peer_exchange():
    """Exchange the BatchNorm statistics of folded tensor-core chains inside the fold kernel, over NVLink peer memory,
    instead of one NCCL all-reduce per BatchNorm: every rank allocates a symmetric buffer pair (torch symmetric memory
    is the plumbing that maps the peers' buffers), and the kernel stores / flags / sums across ranks itself.  Returns
    True when the buffers are mapped; False (reason in ``peer_error``) leaves the all-reduce hook in place."""
    global _peer, peer_error
    if not _enabled or world_size() == 1:
        return False
    from . import _lib
    try:
        import torch.distributed as td
        import torch.distributed._symmetric_memory as sm
        w = world_size()
        if w > _lib.TNF_PEER_MAX:
            raise RuntimeError("more than %d ranks" % _lib.TNF_PEER_MAX)
        dev = torch.device("cuda", torch.cuda.current_device())
        grp = _group if _group is not None else td.group.WORLD
        stats = sm.empty(2 * w * _lib.TNF_PEER_SLOT, dtype=torch.float64, device=dev)
        flags = sm.empty(w, dtype=torch.int64, device=dev)
        stats.zero_()
        flags.zero_()
        hs = sm.rendezvous(stats, grp)
        hf = sm.rendezvous(flags, grp)
        torch.cuda.synchronize()
        td.barrier(group=_group)          # every rank's counters are zero before anyone can signal
        _peer = dict(stats=stats, flags=flags, stats_ptrs=[int(p) for p in hs.buffer_ptrs],
                     flag_ptrs=[int(p) for p in hf.buffer_ptrs], seq=1, handles=(hs, hf))
        peer_error = None
        return True
    except Exception as exc:          # no symmetric memory on this build / topology: keep the NCCL hook
        _peer, peer_error = None, "%s: %s" % (type(exc).__name__, exc)
        return False


def peer_struct(n_exchanges):
    """ctypes ``tnf_peer_t`` for a call that may perform ``n_exchanges`` statistics exchanges (None when peer exchange
    is off).  The sequence counter advances identically on every rank."""
    if _peer is None:
        return None
    from . import _lib
    p = _lib.Peer()
    p.rank, p.world = rank(), world_size()
    for r in range(p.world):
        p.stats[r] = _peer["stats_ptrs"][r]
        p.flags[r] = _peer["flag_ptrs"][r]
    p.seq = _peer["seq"]
    _peer["seq"] += max(1, int(n_exchanges))
    return p


def is_enabled():
    return _enabled


def world_size():
    if not _enabled:
        return 1
    import torch.distributed as td
    return td.get_world_size(_group)


def rank():
    if not _enabled:
        return 0
    import torch.distributed as td
    return td.get_rank(_group)


def shard_range(total, rank_=None, world=None):
    """Contiguous [lo, hi) slice of ``total`` rows owned by a rank (the first
    ``total % world`` ranks hold one extra row)."""
    world = world_size() if world is None else world
    rank_ = rank() if rank_ is None else rank_
    base, rem = divmod(total, world)
    lo = rank_ * base + min(rank_, rem)
    return lo, lo + base + (1 if rank_ < rem else 0)


def allreduce_stats(sums):
    """Sum the float64 ``[sum | sumsq | rows]`` buffer of ``tnf_colstats`` over
    the shards (in place).  The global row count rides in the last slot, so no
    host synchronisation is needed."""
    if not _enabled or world_size() == 1:
        return sums
    import torch.distributed as td
    td.all_reduce(sums, op=td.ReduceOp.SUM, group=_group)
    return sums


def allreduce_sum_(tensors):
    """In-place sum of a list of tensors (gradients) across ranks, one flat bucket."""
    if not _enabled or world_size() == 1:
        return tensors
    import torch.distributed as td
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    td.all_reduce(flat, op=td.ReduceOp.SUM, group=_group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors


def allreduce_mean_scalar(value):
    """Mean of a 0-dim tensor across ranks (loss / mean log-prob)."""
    if not _enabled or world_size() == 1:
        return value
    import torch.distributed as td
    v = value.detach().clone()
    td.all_reduce(v, op=td.ReduceOp.SUM, group=_group)
    return v / world_size()
