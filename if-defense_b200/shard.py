"""Multi-GPU sharding of the restoration job: clouds are independent units, so every reference batch
is cut into `world` contiguous slices, rank r restores slice r of every batch with that batch's B_ref (the losses
are batch means -- SURVEY.md F8 -- so results do not depend on the split), and ONE all_gather of the restored
[n_local, K, 3] blocks at the end puts them back in input order (labels are copied positionally by the caller,
opt_defense.py:336-343).  There is no per-iteration communication.

One process per GPU, `torch.distributed` for the plumbing: backend "nccl" on the GPUs (NVLink/NVSwitch), "gloo" in
the CPU tests of the host logic.  The reference's restoration scripts are single-GPU (ConvONet/command.txt:3); its
multi-GPU launch model -- one rank per GPU, contiguous data split, results concatenated in rank order -- is the one
of its attack scripts (baselines/attack_scripts/targeted_perturb_attack.py:99-174, util/merge_attack_results.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def batch_bounds(n_clouds, batch_size):
    """The reference's batching: range(0, N, batch_size) (opt_defense.py:272) -> [(lo, hi)]."""
    return [(lo, min(lo + batch_size, n_clouds)) for lo in range(0, n_clouds, batch_size)]


def plan(n_clouds, batch_size, world):
    """-> per rank, a list of segments (lo, hi, B_ref): rank r owns clouds lo..hi-1 of a reference batch of B_ref
    clouds.  Slices are contiguous and balanced to within one cloud per batch; empty slices are dropped."""
    out = [[] for _ in range(world)]
    for lo, hi in batch_bounds(n_clouds, batch_size):
        n = hi - lo
        for r in range(world):
            a, b = lo + (n * r) // world, lo + (n * (r + 1)) // world
            if b > a:
                out[r].append((a, b, n))
    return out


def gather_in_order(local, segments_per_rank, n_clouds, rank, world, device=None):
    """local: this rank's restored clouds [n_local, K, 3] (numpy or tensor), concatenated in segment order.
    One all_gather (padded to the largest n_local) -> numpy [n_clouds, K, 3] in input order, on every rank."""
    t = torch.as_tensor(local)
    if device is not None:
        t = t.to(device)
    K = t.shape[1] if t.dim() == 3 else 0
    counts = [sum(b - a for a, b, _ in segs) for segs in segments_per_rank]
    pad = max(counts) if counts else 0
    buf = torch.zeros((pad, K, 3), dtype=torch.float32, device=t.device)
    if counts[rank]:
        buf[:counts[rank]] = t.float()
    if world > 1:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
    else:
        parts = [buf]
    out = np.zeros((n_clouds, K, 3), dtype=np.float32)
    for r, segs in enumerate(segments_per_rank):
        got = parts[r].cpu().numpy()
        off = 0
        for a, b, _ in segs:
            out[a:b] = got[off:off + (b - a)]
            off += b - a
    return out


def restore_sharded(restore_fn, n_clouds, batch_size, rank=None, world=None, device=None, restore_many=None):
    """restore_fn(lo, hi, B_ref) -> restored clouds lo..hi-1 ([hi-lo, K, 3]) with the batch-mean factor 1/B_ref.
    Runs this rank's segments and returns all n_clouds restored clouds in input order (on every rank).
    restore_many (optional): callable(list of (lo, hi, B_ref)) -> list of arrays, used instead of one restore_fn call per
    segment (lets the caller overlap the stages of consecutive segments)."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if world > 1 else 0
    segs = plan(n_clouds, batch_size, world)
    if restore_many is not None:
        mine = [np.asarray(x, dtype=np.float32) for x in restore_many(list(segs[rank]))]
    else:
        mine = [np.asarray(restore_fn(a, b, n), dtype=np.float32) for a, b, n in segs[rank]]
    K = _agree_K(mine[0].shape[1] if mine else None, world)       # a rank with no segment learns K from the others
    local = np.concatenate(mine, axis=0) if mine else np.zeros((0, K, 3), dtype=np.float32)
    return gather_in_order(local, segs, n_clouds, rank, world, device)


def _agree_K(K, world):
    """All ranks agree on K (a rank with no segment learns it from the others): max-reduce."""
    if world == 1:
        return K
    t = torch.tensor([0 if K is None else int(K)], dtype=torch.int64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())
