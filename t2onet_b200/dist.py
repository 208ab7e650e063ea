"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL on GPUs, gloo on CPU).

The path shards by image with NO data-path collective:
  * operator chains / training batches: the batch dimension is split across ranks (DDP does the actor's
    gradient all-reduce; the operators themselves never communicate);
  * planner, image-sharded (default): image i -> rank i mod R, zero traffic during the search, one
    all_gather of the per-image result records at the end;
  * planner, candidate-sharded (optional): candidates of one image are split across ranks and the best
    one is selected with ONE all_reduce(MIN) over packed 64-bit keys  (float_bits(score) << 32 | cand_id).
"""
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def rank_world(group=None):
    if not is_dist():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_indices(n, rank=None, world=None):
    """Indices of the items (images) rank `rank` owns: rank, rank + world, ..."""
    if rank is None or world is None:
        rank, world = rank_world()
    return list(range(rank, n, world))


def gather_records(local, n_total, group=None):
    """local: dict {global index: record} of this rank -> list of n_total records on every rank."""
    rank, world = rank_world(group)
    if world == 1:
        return [local[i] for i in range(n_total)]
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    merged = {}
    for p in parts:
        merged.update(p)
    missing = [i for i in range(n_total) if i not in merged]
    if missing:
        raise RuntimeError('planner records missing for items %s' % missing[:8])
    return [merged[i] for i in range(n_total)]


def pack_score_keys(scores, ids):
    """(score >= 0 fp32, id < 2^31) -> int64 keys whose integer order is (score, id) order."""
    bits = scores.detach().float().contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return (bits << 32) | ids.to(torch.int64)


def unpack_score_keys(keys):
    ids = (keys & 0xFFFFFFFF).to(torch.int64)
    scores = (keys >> 32).to(torch.int32).view(torch.float32)
    return scores, ids


def best_candidate(scores, ids, group=None):
    """Global argmin over candidates sharded across ranks.

    scores (M, C_local) non-negative fp32, ids (M, C_local) global candidate ids: one row per (image, state)
    in flight.  Returns (best_score (M,), best_id (M,)) identical on every rank, using a single
    all_reduce(MIN) of 8 bytes per row (latency-bound, NVLink/NVSwitch via NCCL on GPUs)."""
    keys = pack_score_keys(scores, ids)
    local_best = keys.min(dim=1).values if keys.dim() == 2 else keys
    if is_dist() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_best, op=dist.ReduceOp.MIN, group=group)
    return unpack_score_keys(local_best)


def plan_dataset(pairs, executor, plan_fn, group=None):
    """Image-sharded planning: `pairs` is an indexable of (I_0, I_gt); rank r plans items r, r+R, ...
    with plan_fn(I_0, I_gt, executor) -> JSON-able record; every rank returns the full, ordered list."""
    rank, world = rank_world(group)
    local = {}
    for i in shard_indices(len(pairs), rank, world):
        I_0, I_gt = pairs[i]
        local[i] = plan_fn(I_0, I_gt, executor)
    return gather_records(local, len(pairs), group)


def plan_dataset_batched(pairs, executor, batch_fn, batch=64, group=None):
    """Image-sharded planning in lock-step batches: rank r takes items r, r+R, ... and hands them to
    batch_fn(indices, [pairs[i] for i in indices], executor) -> list of JSON-able records `batch` at a time
    (e.g. stacking the pairs for planner.beam_search_batch, whose device-resident fits want many pairs in flight).
    No communication during the search; every rank returns the full, ordered list."""
    rank, world = rank_world(group)
    mine = list(shard_indices(len(pairs), rank, world))
    local = {}
    for c0 in range(0, len(mine), batch):
        idx = mine[c0:c0 + batch]
        recs = batch_fn(idx, [pairs[i] for i in idx], executor)
        assert len(recs) == len(idx)
        local.update(zip(idx, recs))
    return gather_records(local, len(pairs), group)
