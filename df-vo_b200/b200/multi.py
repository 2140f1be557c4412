"""Multi-GPU plumbing of the tracking path: one process per GPU, one image sequence per process.

DF-VO's per-frame path has no cross-sequence data dependency (dfvo.py:358-403 walks one sequence), so the path
shards over *sequences*: rank r owns sequences r, r+world, ... and runs them with the weights every rank
received once at start-up.  That broadcast is the only collective; there is no data-path exchange, hence no
fused compute+collective kernel to write.  Timing is the max over ranks.

Works on any torch.distributed backend: NCCL on the GPU box, gloo in the CPU test-suite
(tests/test_multi_gloo.py).
"""
import numpy as np


def weight_layout(dicts):
    """Deterministic (dict index, key, shape, offset) table over the array-valued entries of the state dicts."""
    table, off = [], 0
    for di, d in enumerate(dicts):
        for k in sorted(d):
            v = d[k]
            if hasattr(v, "shape") and getattr(v, "dtype", None) is not None:
                n = int(np.prod(v.shape)) if len(v.shape) else 1
                table.append((di, k, tuple(v.shape), off, n))
                off += n
    return table, off


def pack_weights(dicts):
    table, total = weight_layout(dicts)
    flat = np.empty(total, np.float32)
    for di, k, shape, off, n in table:
        flat[off:off + n] = np.asarray(dicts[di][k], np.float32).reshape(-1)
    return flat, table


def unpack_weights(flat, table, dicts):
    for di, k, shape, off, n in table:
        dicts[di][k] = np.array(flat[off:off + n], np.float32).reshape(shape)
    return dicts


def broadcast_weights(dicts, src=0, device=None):
    """Every rank passes state dicts of identical structure (keys + shapes); on return all hold rank `src`'s
    values.  One flat fp32 buffer, one collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dicts
    flat, table = pack_weights(dicts)
    t = torch.from_numpy(flat)
    if device is not None:
        t = t.to(device)
    if dist.get_rank() != src:
        t.zero_()
    dist.broadcast(t, src=src)
    return unpack_weights(t.cpu().numpy(), table, dicts)


def rank_sequences(n_sequences, rank, world):
    """Round-robin ownership of sequences (stripes of the dataset); every sequence has exactly one owner."""
    return list(range(rank, n_sequences, world))


def max_over_ranks(ms, device=None):
    """A timed region's duration is the slowest rank's."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_trajectories(poses, dst=0):
    """Collect each rank's {sequence: {frame: 4x4}} on `dst` (evaluation / saving happens on one rank)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [poses]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(poses, out, dst=dst)
    return out
