"""Multi-GPU sharding of the two hot paths (one process per GPU, torch.distributed).

Mapping records (path 2) and query fragments (path 1) are independent units
(src/align/include/computeAlignments.hpp:398-435, src/map/include/computeMap.hpp:565-599), so a batch is
partitioned across ranks with NO data-path collective; NCCL (or gloo on CPU) is used only to gather the
variable-length results on rank 0, which owns filtering / PAF output like the reference's single writer."""
import heapq


def record_cost(plen, tlen, est_identity=None):
    """Cost model of one mapping record for load balancing: biWFA work grows with score^2 and the score
    with (1 - identity) * length; the mapping PAF's id:f: estimate is a free predictor."""
    d = 0.05 if est_identity is None else max(0.002, 1.0 - float(est_identity))
    s = d * max(plen, tlen) * 6.4 + abs(plen - tlen) + 50.0
    return s * s


def partition(costs, world):
    """Longest-processing-time-first partition of unit indices over `world` ranks.
    Returns a list (per rank) of index lists, each in ascending index order; deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(world)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    for s in shards:
        s.sort()
    return shards


def gather_results(local_indices, local_results, n_total, group=None, dst=0):
    """Gather per-unit results of all ranks on `dst` in the original unit order.
    Works with any initialised torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    payload = (list(local_indices), list(local_results))
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * n_total
    for idx, res in gathered:
        for i, r in zip(idx, res):
            out[i] = r
    return out


def align_sharded(pairs, align_fn, identities=None, group=None):
    """Partition `pairs` over the ranks of the process group by cost, align the local shard with
    align_fn(list_of_pairs) -> list_of_results, gather on rank 0. Returns the full list on rank 0, None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    costs = [record_cost(len(p), len(t), None if identities is None else identities[i]) for i, (p, t) in enumerate(pairs)]
    mine = partition(costs, world)[rank]
    res = align_fn([pairs[i] for i in mine]) if mine else []
    return gather_results(mine, res, len(pairs), group=group)
