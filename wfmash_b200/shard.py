"""Multi-GPU sharding of the two hot paths (one process per GPU, torch.distributed).

Mapping records (path 2) and query fragments (path 1) are independent units
(src/align/include/computeAlignments.hpp:398-435, src/map/include/computeMap.hpp:565-599), so a job is
partitioned across ranks with NO data-path collective; NCCL (or gloo on CPU) only exchanges the variable-length texts.

Two flavours:
  * job_sharded / partition_queries / partition_rows / allgather_bytes — over the C++ phases of the library (wfb_map_phase_subset,
    wfb_align_phase): what `bench.py --gpus N` runs and what a caller of the C ABI would do;
  * map_sharded / align_paf_sharded / align_sharded — the round-1 helpers over the Python mirror of the phases (pipeline.py), kept for
    record-level work (rank 0 receives per-unit results in unit order)."""
import heapq
import time


# ---------------------------------------------------------------------------------------------------------------------------
# One job over the ranks through the C++ phases
# ---------------------------------------------------------------------------------------------------------------------------
def partition_queries(queries, w, world):
    """Whole queries per rank (the chain / filter stage needs every fragment of a query, computeMap.hpp:635-667), longest first onto
    the lightest rank. -> {query name: rank}"""
    load, owner = [0] * world, {}
    for n, s in sorted(queries, key=lambda x: -len(x[1])):
        r = min(range(world), key=lambda i: load[i])
        owner[n] = r
        load[r] += len(s) if len(s) >= w else 0
    return owner


def partition_rows(rows, world):
    """Mapping rows over the ranks by expected cost ((1 - identity) * length)^2, heaviest first onto the lightest rank (LPT).
    -> [rank of row i]"""
    cost = []
    for ln in rows:
        f = ln.split(b"\t")
        ident = 0.95
        for x in f[12:]:
            if x.startswith(b"id:f:"):
                ident = float(x[5:])
        d = max(1.0 - ident, 0.002)
        cost.append((d * (int(f[3]) - int(f[2]))) ** 2 + 1e4)
    load, owner = [0.0] * world, [0] * len(rows)
    for i in sorted(range(len(rows)), key=lambda j: -cost[j]):
        r = min(range(world), key=lambda k: load[k])
        owner[i] = r
        load[r] += cost[i]
    return owner


def allgather_bytes(data: bytes, dev, world, group=None):
    """All-gather of one byte string per rank (sizes first, then the padded payloads): NCCL on device tensors, gloo on CPU tensors.
    -> ([bytes of rank 0, ...], total bytes)"""
    import torch
    import torch.distributed as dist
    n = torch.tensor([len(data)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if data:
        buf[: len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    out = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return [bytes(out[r][: sizes[r]].cpu().numpy()) for r in range(world)], sum(sizes)


def job_sharded(wb, aligner, targets, queries, map_params, window_length, device, tensor_device, rank, world, batch_records=0, group=None,
                my_queries=None):
    """ONE `wfmash targets queries` job over `world` ranks through the C++ phases; every rank returns the whole job's texts.
    Mapping: every rank builds the (replicated) index and maps the queries it owns (ids, PanSN groups, ANI estimate and fragment order
    come from all queries: wfb_map_phase_subset); the ranks' rows are all-gathered and put back in the single-process order (queries in
    input order, each query's rows as its rank wrote them). Alignment: rows partitioned by expected cost, each rank aligns its share,
    the PAF bytes are all-gathered (concatenated in rank order: the reference's own record order is its worker threads' completion
    order, computeAlignments.hpp:535-542). -> dict(mapping_paf, paf, map_stats, align_stats, my_records, timings...)"""
    r = {}
    t0 = time.perf_counter()
    same = queries is targets
    if world == 1:
        mp, mst = wb.map_phase(targets, queries, map_params, device)
    else:
        if my_queries is None:
            owner = partition_queries(queries, window_length, world)
            my_queries = [(n, s) for n, s in queries if owner[n] == rank]
        mp, mst = wb.map_phase(targets, my_queries, map_params, device, all_queries=queries)
    r["t_map"] = time.perf_counter() - t0
    r["gather_bytes"] = 0
    if world > 1:
        t1 = time.perf_counter()
        parts, nb = allgather_bytes(mp, tensor_device, world, group)
        per_query = {}
        for part in parts:
            for ln in part.split(b"\n"):
                if ln:
                    per_query.setdefault(ln.split(b"\t", 1)[0], []).append(ln)
        rows = [ln for n, _ in queries for ln in per_query.get(n.encode() if isinstance(n, str) else n, [])]
        owner = partition_rows(rows, world)
        my_rows = b"".join(rows[i] + b"\n" for i in range(len(rows)) if owner[i] == rank)
        r["t_exchange_rows"] = time.perf_counter() - t1
        r["gather_bytes"] += nb
        full_mp = b"".join(x + b"\n" for x in rows)
    else:
        my_rows, full_mp = mp, mp
    t2 = time.perf_counter()
    paf, ast = wb.align_phase(aligner, my_rows, targets, queries if not same else targets, window_length=window_length, batch_records=batch_records)
    r["t_align"] = time.perf_counter() - t2
    if world > 1:
        t3 = time.perf_counter()
        parts, nb = allgather_bytes(paf, tensor_device, world, group)
        r["t_gather_paf"] = time.perf_counter() - t3
        r["gather_bytes"] += nb
        paf_all = b"".join(parts)
    else:
        paf_all = paf
    r["t_total"] = time.perf_counter() - t0
    r.update(mst=mst, ast=ast, mapping_paf=full_mp, paf=paf_all, my_records=int(ast.records))
    return r


# ---------------------------------------------------------------------------------------------------------------------------
# Record-level helpers over the Python mirror of the phases (round 1)
# ---------------------------------------------------------------------------------------------------------------------------


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


def map_sharded(targets, queries, params=None, device=0, group=None, map_fn=None):
    """Mapping phase over the ranks of a process group: every rank holds a replica of the index (C5 needs ~64 GB of
    180 GB) and maps the queries of its partition — whole queries, because the chain / filter stage needs all fragments of a
    query (computeMap.hpp:635-667) — balanced by length. The mapping PAF of all ranks is gathered on rank 0 in query order
    (None elsewhere). The one-to-one mode needs every query's survivors for its final pass and is not sharded."""
    import torch.distributed as dist
    from wfmash_b200 import pipeline
    if map_fn is None:
        def map_fn(only):
            return pipeline.map(targets, queries, params, device, only_queries=only).paf
    P = params or pipeline.Params()
    if P.filter is not None and P.filter.filter_mode == 2:
        raise NotImplementedError("one-to-one filtering runs a final pass over all queries (computeMap.hpp:788-850): map on one rank")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = partition([float(len(s)) for _, s in queries], world)[rank]
    paf = map_fn({queries[i][0] for i in mine}) if mine else b""
    by_query = {}
    for line in paf.split(b"\n"):
        if line:
            by_query.setdefault(line.split(b"\t", 1)[0].decode(), []).append(line)
    res = gather_results(mine, [b"".join(l + b"\n" for l in by_query.get(queries[i][0], [])) for i in mine], len(queries), group=group)
    return None if res is None else b"".join(res)


def align_paf_sharded(mapping_paf, targets, queries, params=None, device=0, group=None, align_fn=None):
    """Alignment phase over the ranks of a process group: the rows of the mapping PAF are partitioned by estimated cost
    (record_cost from the row's spans and id:f: tag), every rank aligns its rows, rank 0 receives the records in row order
    (the reference's single writer, computeAlignments.hpp:535-542). No collective touches the kernels."""
    import torch.distributed as dist
    from wfmash_b200 import pipeline
    if align_fn is None:
        def align_fn(text, per_row=True):
            return pipeline.align(text, targets, queries, params, device, per_row=per_row)[0]
    rows = [ln for ln in mapping_paf.split(b"\n") if ln]
    costs = []
    for ln in rows:
        f = ln.split(b"\t")
        try:
            ident = float(f[12].rsplit(b":", 1)[-1])
            costs.append(record_cost(int(f[8]) - int(f[7]), int(f[3]) - int(f[2]), ident))
        except (IndexError, ValueError):
            costs.append(1.0)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = partition(costs, world)[rank]
    # one call per row keeps the row -> record association without parsing the output (a row may produce no record)
    out = align_fn(b"".join(rows[i] + b"\n" for i in mine), per_row=True) if mine else []
    res = gather_results(mine, out, len(rows), group=group)
    return None if res is None else b"".join(res)
