"""Multi-process host logic on CPU: world_size-2 gloo run of the record sharding + result gather
(the N>1 path of bench.py / align_sharded with a stand-in for the GPU call)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from wfmash_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = [(b"A" * (100 + 37 * i), b"C" * (90 + 41 * i)) for i in range(23)]
    seen = []

    def fake_align(local):  # stands in for Aligner.align_end2end_batch on this rank's GPU
        seen.extend(local)
        return [(len(p), len(t), rank) for p, t in local]

    out = shard.align_sharded(pairs, fake_align)
    # the two phases of the pipeline sharded the same way (stand-ins for the GPU calls; the real ones are pipeline.map / align)
    queries = [(f"s{i}#1#c", b"A" * (1000 * (i % 5 + 1))) for i in range(9)]

    def fake_map(only):  # two mapping rows per query of this rank's partition, one of them unparsable
        return b"".join(b"%s\t%d\t0\t%d\t+\tt\t99999\t10\t%d\t5\t100\t30\tid:f:0.9%d\tkc:f:1\n%s\tgarbage\n"
                        % (n.encode(), len(s), len(s), 10 + len(s), rank, n.encode()) for n, s in queries if n in only)

    mp = shard.map_sharded(None, queries, None, map_fn=fake_map)
    box = [mp]
    dist.broadcast_object_list(box, src=0)

    def fake_align(text, per_row=True):  # one record per parsable row, none for the garbage rows
        return [(b"aligned:" + ln.split(b"\t")[0] + b":%d\n" % rank) if b"garbage" not in ln else b"" for ln in text.split(b"\n") if ln]

    paf = shard.align_paf_sharded(box[0], None, queries, align_fn=fake_align)
    if rank == 0:
        q.put((out, len(seen), mp, paf))
    else:
        q.put((None, len(seen), None, None))
    dist.destroy_process_group()


def test_partition_is_balanced_and_complete():
    sys.path.insert(0, ROOT)
    from wfmash_b200 import shard
    costs = [shard.record_cost(1000 * (i % 17 + 1), 900 * (i % 13 + 1), 0.9 + 0.005 * (i % 10)) for i in range(200)]
    for world in (1, 2, 4, 8):
        sh = shard.partition(costs, world)
        assert sorted(i for s in sh for i in s) == list(range(200))
        loads = [sum(costs[i] for i in s) for s in sh]
        assert max(loads) <= 1.25 * (sum(costs) / world) + max(costs)
    assert shard.partition(costs, 4) == shard.partition(costs, 4)  # deterministic


def _job_worker(rank, world, port, q):
    """wfmash_b200.shard.job_sharded (the flow bench.py --gpus N runs) with stand-ins for the two C-ABI phases: gloo all-gathers of byte
    tensors, rows re-assembled in the single-process order, rows partitioned by expected cost, PAF gathered on every rank."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import types
    import torch
    import torch.distributed as dist
    from wfmash_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    queries = [(f"s{i}#1#c", b"A" * (1000 * (i % 5 + 1))) for i in range(11)] + [("short#1#c", b"A" * 10)]
    calls = {}

    class FakeWb:  # stands in for the ctypes mirror: two rows per mapped query (identity depends on the query), one PAF line per row
        @staticmethod
        def map_phase(targets, mine, params, device, all_queries=None):
            assert all_queries is queries and {n for n, _ in mine} <= {n for n, _ in queries}
            calls["mapped"] = [n for n, _ in mine]
            txt = b"".join(b"%s\t%d\t0\t%d\t+\tt%d\t99999\t10\t%d\t5\t100\t30\tid:f:0.%d\tkc:f:1\n" % (n.encode(), len(s), len(s), j, 10 + len(s), 80 + len(s) // 1000)
                           for n, s in mine if len(s) >= 1000 for j in range(2))
            return txt, types.SimpleNamespace(rank=rank)

        @staticmethod
        def align_phase(aligner, rows, targets, qs, window_length=1000, batch_records=0):
            lines = [ln for ln in rows.split(b"\n") if ln]
            calls["aligned"] = len(lines)
            return b"".join(b"paf:" + ln.split(b"\t")[0] + b":" + ln.split(b"\t")[5] + b":%d\n" % rank for ln in lines), types.SimpleNamespace(records=len(lines))

    r = shard.job_sharded(FakeWb, None, queries, queries, None, 1000, 0, torch.device("cpu"), rank, world)
    q.put((rank, r["mapping_paf"], r["paf"], calls["mapped"], calls["aligned"], r["my_records"], r["gather_bytes"]))
    dist.destroy_process_group()


def test_two_rank_gloo_job_through_the_phase_level_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_job_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, mp0, paf0, mapped0, n0, rec0, gb0), (_, mp1, paf1, mapped1, n1, rec1, gb1) = got
    assert mp0 == mp1 and paf0 == paf1 and gb0 == gb1 > 0                     # every rank holds the whole job's texts
    assert sorted(mapped0 + mapped1) == sorted([f"s{i}#1#c" for i in range(11)] + ["short#1#c"]) and mapped0 and mapped1   # queries split, none twice
    rows = [ln for ln in mp0.split(b"\n") if ln]
    assert [ln.split(b"\t")[0].decode() for ln in rows] == [f"s{i}#1#c" for i in range(11) for _ in range(2)]   # single-process order restored
    assert n0 + n1 == len(rows) == rec0 + rec1 and n0 > 0 and n1 > 0            # rows split over both ranks, each aligned once
    recs = [ln for ln in paf0.split(b"\n") if ln]
    assert sorted(b":".join(r.split(b":")[1:3]) for r in recs) == sorted(ln.split(b"\t")[0] + b":" + ln.split(b"\t")[5] for ln in rows)
    assert {r.split(b":")[3] for r in recs} == {b"0", b"1"}


def test_two_rank_gloo_shard_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = [g for g in got if g[0] is not None]
    assert len(full) == 1
    out, _, mp, paf = full[0]
    rows = [ln for ln in mp.split(b"\n") if ln]
    assert [ln.split(b"\t")[0].decode() for ln in rows] == [f"s{i}#1#c" for i in range(9) for _ in range(2)]   # query order restored, rows kept together
    assert {ln.split(b"\t")[12] for ln in rows if b"garbage" not in ln} == {b"id:f:0.90", b"id:f:0.91"}        # both ranks mapped
    recs = [ln for ln in paf.split(b"\n") if ln]
    assert [r.split(b":")[1].decode() for r in recs] == [f"s{i}#1#c" for i in range(9)] and {r.split(b":")[2] for r in recs} == {b"0", b"1"}
    assert len(out) == 23 and all(o is not None for o in out)
    assert [(o[0], o[1]) for o in out] == [(100 + 37 * i, 90 + 41 * i) for i in range(23)]
    assert {o[2] for o in out} == {0, 1}            # both ranks did work
    assert sum(g[1] for g in got) == 23             # every record aligned exactly once
