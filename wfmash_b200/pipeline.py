"""The two phases of `wfmash` over sequences that are already in host memory (no CLI, no FASTA / index files): the
host-side mirror of what src/interface/main.cpp drives —

  map():    skch::Map (src/map/include/computeMap.hpp:300-860): sequence ids + PanSN groups, index build, query
            fragments, L1 + L2 on the GPU, per-query chain merge + filters on the host, mapping PAF text;
  align():  align::Aligner (src/align/include/computeAlignments.hpp:318-720): mapping PAF rows -> padded records ->
            strand-corrected, N-masked slices -> batched do_biwfa_alignment on the GPU -> alignment PAF text;
  wfmash(): both, handing the mapping PAF text from one to the other exactly like the reference does through its
            temporary file (main.cpp: mashmapPafFile).

Every compute step is a call into libwfmash_b200.so (wfmash_b200/__init__.py); nothing here computes on the CPU except
slicing / reverse-complementing the record sequences, which the reference also does on the host
(computeAlignments.hpp:664-686)."""
import dataclasses

import numpy as np

import wfmash_b200 as wb

_UPPER_VALID = np.full(256, ord("N"), dtype=np.uint8)  # CommonFunc::makeUpperCaseAndValidDNA, commonFunc.hpp:132-142
for _c in b"ACGT":
    _UPPER_VALID[_c] = _c
    _UPPER_VALID[_c + 32] = _c
_COMPLEMENT = np.arange(256, dtype=np.uint8)           # CommonFunc::reverseComplement, commonFunc.hpp:74-83: only A/C/G/T swap
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMPLEMENT[_a] = _b


@dataclasses.dataclass
class Params:
    """The skch::Parameters / align::Parameters fields the two hot paths read, with the CLI defaults of
    src/interface/parse_args.hpp (except percentage_identity: the CLI default estimates it from the data (ANI);
    here it is explicit, like `-p`)."""
    kmer_size: int = 15
    window_length: int = 1000              # -w (segment length)
    percentage_identity: float = 0.90      # -p 90; None = the CLI default `-p ani50-2` (estimate_identity() below)
    ani_percentile: int = 50
    ani_adjustment: float = -2.0
    sketch_size: int = 0                   # -s; 0 = derive from identity / window / k
    minimum_hits: int = -1                 # -H; -1 = auto
    max_kmer_freq: float = 0.0002          # -F
    index_threads: int = 1
    skip_self: bool = True                 # no -X
    skip_prefix: bool = True               # -Y '#'
    prefix_delim: str = "#"
    lower_triangular: bool = False         # -L
    stage1_top_ani_filter: bool = True
    hg_numerator: float = 1.0
    ani_diff: float = 0.0
    ani_diff_conf: float = 0.999
    keep_low_pct_id: bool = True
    filter: wb.FilterParams = None         # chain / filter stage; None = CLI defaults for this window length
    target_padding: int = -1               # -E; -1 = min(window_length, 5000)
    query_padding: int = -1                # -U; -1 = min(window_length, 5000)
    min_identity: float = 0.0              # alignment output filters (parse_args.hpp:566-584)
    min_alignment_length: int = 32
    min_block_identity: float = 0.1
    disable_chain_patching: bool = False
    term_group: int = 8                    # ends-free termination order of the reference build to reproduce (8 = AVX2)

    def resolved(self):
        p = dataclasses.replace(self)
        if p.percentage_identity is None:
            raise ValueError("percentage_identity is None: call pipeline.auto_identity(targets, queries, params) first (the CLI does this before mapping, main.cpp:75-134)")
        if p.sketch_size <= 0:
            p.sketch_size = wb.sketch_size(p.percentage_identity, p.window_length, p.kmer_size)
        if p.filter is None:
            p.filter = wb.FilterParams(window_length=p.window_length, percentage_identity=p.percentage_identity, skip_prefix=int(p.skip_prefix))
        if p.target_padding < 0:
            p.target_padding = min(p.window_length, 5000)
        if p.query_padding < 0:
            p.query_padding = min(p.window_length, 5000)
        return p


class SequenceIds:
    """skch::SequenceIdManager (src/map/include/sequenceIds.hpp:284-441): ids in first-seen order, targets before queries;
    group = everything before the LAST prefix delimiter, numbered in sorted-name order."""

    def __init__(self, targets, queries, prefix_delim="#"):
        self.names, self.lengths, self.id_of = [], [], {}
        for name, seq in list(targets) + list(queries):
            if name not in self.id_of:
                self.id_of[name] = len(self.names)
                self.names.append(name)
                self.lengths.append(len(seq))
        keys, self.group = {}, [0] * len(self.names)
        for name in sorted(self.names):
            pos = name.rfind(prefix_delim) if prefix_delim else -1
            key = name[:pos] if pos >= 0 else name
            if key not in keys:
                keys[key] = len(keys) + 1
            self.group[self.id_of[name]] = keys[key]


def estimate_identity(targets, queries, params: Params = None, device: int = 0):
    """skch::Stat::estimate_identity_for_groups (map_stats.hpp:325-822): k = 21, 4096-hash MinHash per PanSN group and role,
    all query-role groups against all target-role groups of another id -> (identity, stats). The sketches come from the GPU."""
    P = params or Params()
    ids = SequenceIds(targets, queries, P.prefix_delim if P.skip_prefix else "")
    gids = sorted(set(ids.group))
    dense = {g: i for i, g in enumerate(gids)}
    same = [n for n, _ in targets] == [n for n, _ in queries]

    def role(seqs):
        sk, cnt, st = wb.ani_group_sketches([s for _, s in seqs], [dense[ids.group[ids.id_of[n]]] for n, _ in seqs], len(gids), 21, 4096, device)
        present = sorted({dense[ids.group[ids.id_of[n]]] for n, _ in seqs})
        return sk[present], cnt[present], [gids[i] for i in present], st

    q = role(queries)
    t = q if same else role(targets)
    ident, ncmp = wb.ani_estimate_identity(q[0], q[1], q[2], t[0], t[1], t[2], 21, P.ani_percentile, P.ani_adjustment)
    return ident, {"comparisons": ncmp, "hash_kernel_ms": q[3].hash_kernel_ms + (0 if same else t[3].hash_kernel_ms), "valid_kmers": int(q[3].valid_kmers)}


def auto_identity(targets, queries, params: Params = None, device: int = 0) -> Params:
    """main.cpp:75-134: adopt the estimated identity and, unless -s was given, re-derive the sketch size from it."""
    P = dataclasses.replace(params or Params())
    P.percentage_identity, _ = estimate_identity(targets, queries, P, device)
    if params is None or params.sketch_size <= 0:
        P.sketch_size = min(wb.sketch_size(P.percentage_identity, P.window_length, P.kmer_size), P.window_length)
    return P


def one_thread_fragment_order(n_fragments, queue_capacity: int = 255):
    """The order in which a ONE-worker taskflow executor (`wfmash -t 1`) runs the fragment tasks of every query, which is the order
    their results reach Map::filterSubsetMappings and therefore decides the ch:Z: tags (computeMap.hpp:532-640).

    n_fragments[q] = number of fragment tasks of the q-th entry of querySequenceNames (0 for sequences shorter than the segment
    length: they still get a query task). Restates the executor's scheduling (src/common/taskflow/core/executor.hpp:1271-1310,
    1452-1493; tsq.hpp:436-484, TF_DEFAULT_BOUNDED_TASK_QUEUE_LOG_SIZE 8 => 255 usable slots): scheduled nodes go to the worker's
    bounded LIFO queue, overflow to the unbounded FIFO free list; a worker waiting in Subflow::join keeps popping its own queue
    (which may hold OTHER queries' tasks: they run nested) and takes from the free list only when the queue is empty.
    -> list of per-query lists of fragment indices."""
    order = [[] for _ in n_fragments]
    remaining = list(n_fragments)
    stack, fifo, head, waiting = [], [], 0, []

    def schedule(node):
        if len(stack) < queue_capacity:
            stack.append(node)
        else:
            fifo.append(node)

    for q in range(len(n_fragments)):
        schedule((q, -1))
    while True:
        if waiting and remaining[waiting[-1]] == 0:
            waiting.pop()
            continue
        if stack:
            q, i = stack.pop()
        elif head < len(fifo):
            q, i = fifo[head]
            head += 1
        else:
            break
        if i < 0:  # a query task: emplaces its fragments, then joins
            for j in range(n_fragments[q]):
                schedule((q, j))
            waiting.append(q)
        else:
            order[q].append(i)
            remaining[q] -= 1
    return order


@dataclasses.dataclass
class MapResult:
    paf: bytes                 # the `wfmash -m` text
    mappings: np.ndarray       # MAPPING_DTYPE, grouped by query
    chain_info: np.ndarray
    query_offset: np.ndarray   # CSR over the queries that were mapped
    query_names: list
    stats: dict


def map(targets, queries, params: Params = None, device: int = 0, index=None, only_queries=None) -> MapResult:  # noqa: A001 - the reference's phase name
    """Mapping phase. targets / queries: [(name, bytes)]. `index` reuses a wb.Index built for the same targets.
    only_queries: names of the queries this call maps (multi-GPU sharding by query); ids and groups still come from ALL sequences."""
    P = (params or Params()).resolved()
    k, w, s = P.kmer_size, P.window_length, P.sketch_size
    ids = SequenceIds(targets, queries, P.prefix_delim if P.skip_prefix else "")
    own_index = index is None
    if own_index:
        index = wb.Index([seq for _, seq in targets], [ids.id_of[n] for n, _ in targets], k, w, s, max_kmer_freq=P.max_kmer_freq,
                         index_threads=P.index_threads, device=device)
    min_hits = max(P.minimum_hits, wb.estimate_minimum_hits_relaxed(s, k, P.percentage_identity))  # Map::cached_minimum_hits, computeMap.hpp:160
    cutoffs = wb.sketch_cutoffs(s, k, P.ani_diff, P.ani_diff_conf, P.stage1_top_ani_filter)
    s1 = wb.stage1_min_hits(k, s, P.hg_numerator, P.ani_diff) if P.stage1_top_ani_filter else None
    shared = wb.l2_min_shared_relaxed(P.percentage_identity, k, s) if P.keep_low_pct_id else wb.l2_min_shared(P.percentage_identity, k, s)

    # query fragments (computeMap.hpp:560-630): floor(len / w) window-sized pieces + one overlapping tail piece
    mapped = [(n, seq) for n, seq in queries if len(seq) >= w and (only_queries is None or n in only_queries)]
    blob = b"".join(seq for _, seq in mapped)
    frags, fqs, frag_index, q_frag = [], [], [], [0]
    base = 0
    for name, seq in mapped:
        qid = ids.id_of[name]
        nfull = len(seq) // w
        for i in range(nfull):
            frags.append((base + i * w, w, qid)); fqs.append((qid, ids.group[qid])); frag_index.append(i)
        if len(seq) % w != 0:
            frags.append((base + len(seq) - w, w, qid)); fqs.append((qid, ids.group[qid])); frag_index.append(nfull)
        q_frag.append(len(frags))
        base += len(seq)
    ref_len = np.array(ids.lengths, dtype=np.int64)
    groups = np.array(ids.group, dtype=np.int32)
    stats = {"fragments": len(frags), "sketch_size": s, "minimum_hits": min_hits}
    if not frags:
        if own_index:
            index.close()
        return MapResult(b"", np.zeros(0, wb.MAPPING_DTYPE), np.zeros(0, wb.CHAIN_INFO_DTYPE), np.zeros(1, np.int64), [], stats)
    r = index.map_fragments(blob, np.array(frags, dtype=wb.FRAG_DTYPE), np.array(fqs, dtype=wb.FRAG_QUERY_DTYPE), min_hits, cutoffs, groups,
                            skip_self=P.skip_self, skip_prefix=P.skip_prefix, lower_triangular=P.lower_triangular, stage1_min_hits=s1,
                            l2_min_shared=shared)
    if (r["status"] != 0).any():
        raise wb.WfbError("a fragment exceeded an internal capacity of the mapping kernels")
    stats.update(l1_kernel_ms=r["l1_kernel_ms"], l2_kernel_ms=r["l2_kernel_ms"], l1_loci=int(r["n_l1_loci"]), l2_mappings=int(len(r["mappings"])))
    if own_index:
        stats["index"] = {"kept_minmers": int(index.stats.kept_minmers), "stream_kernel_ms": index.stats.minmer.stream_kernel_ms}
        index.close()
    fi = np.array(frag_index, dtype=np.int32)
    off = r["offset"]
    # every query of the run has a task in the reference's executor, also the ones that map nothing (or that another rank maps)
    nfr = {n: (len(seq) // w + (1 if len(seq) % w else 0) if len(seq) >= w else 0) for n, seq in queries}
    all_order = one_thread_fragment_order([nfr[n] for n, _ in queries])
    order_of = {n: o for (n, _), o in zip(queries, all_order)}
    frag_order = [order_of[n] for n, _ in mapped]
    per_query, q_off = [], [0]
    for qi, (name, seq) in enumerate(mapped):
        # The order in which the fragments' results reach the chain merge decides the ch:Z: tags; the reference appends them as its
        # fragment tasks finish (computeMap.hpp:590-597). Its only reproducible schedule is the one-thread run: use that order
        # (one_thread_fragment_order), so the text equals `wfmash -m -t 1`.
        parts = [r["mappings"][off[q_frag[qi] + f]: off[q_frag[qi] + f + 1]] for f in frag_order[qi]]
        l2 = np.concatenate(parts) if parts else r["mappings"][:0]
        per_query.append(wb.l2_to_query_mappings(l2, fi, w, len(seq), ref_len))
        q_off.append(q_off[-1] + len(l2))
    allm = np.concatenate(per_query) if per_query else np.zeros(0, wb.MAPPING_DTYPE)
    qlen = np.array([len(seq) for _, seq in mapped], dtype=np.int64)
    out, info, oo = wb.filter_mappings_batch(P.filter, allm, np.array(q_off, dtype=np.int64), qlen, ref_len, ref_group=groups)
    chain = info
    if P.filter.filter_mode == wb.FILTER_ONETOONE:  # computeMap.hpp:788-850: second pass over the reference axis, default chain tags
        kept, owner = wb.one_to_one_filter(P.filter, out, oo, ref_len, ref_group=groups)
        out, oo, chain = kept, np.searchsorted(owner, np.arange(len(mapped) + 1)).astype(np.int64), None
    text = []
    for qi, (name, seq) in enumerate(mapped):
        c = chain[oo[qi]: oo[qi + 1]] if chain is not None else None
        text.append(wb.mapping_paf_format(P.filter, out[oo[qi]: oo[qi + 1]], c, name, len(seq), ids.names, ref_len))
    stats["mappings"] = int(len(out))
    return MapResult(b"".join(text), out, chain if chain is not None else np.zeros(0, wb.CHAIN_INFO_DTYPE), oo, [n for n, _ in mapped], stats)


def records_from_paf(paf: bytes, targets, queries, params: Params = None):
    """Aligner::processMappingRecord up to the do_biwfa_alignment call (computeAlignments.hpp:452-686): one record dict
    (the argument list of do_biwfa_alignment, wflign.hpp:37-62) per parsable line; unparsable lines are skipped like
    the reference skips them (computeAlignments.hpp:368-372)."""
    P = (params or Params()).resolved()
    tseq, qseq = dict(targets), dict(queries)
    recs = []
    for row_index, line in enumerate(ln for ln in paf.split(b"\n") if ln):
        try:
            row, qn, tn = wb.mapping_paf_parse(line, P.target_padding, P.query_padding, P.window_length * 128)
        except wb.WfbError:
            continue
        if tn not in tseq or qn not in qseq:
            continue  # "sequence not found": the reference reports and drops the record, and so does wfb_align_phase
        # createSeqRecord fetches through faidx, which clamps a range to the sequence (src/common/faigz.h:432-438): merged
        # mappings may end beyond the query (blockLength is the larger of the two spans); an empty fetch drops the record
        t = _UPPER_VALID[np.frombuffer(tseq[tn], dtype=np.uint8)[row.r_start: row.r_end]]
        q = _UPPER_VALID[np.frombuffer(qseq[qn], dtype=np.uint8)[row.q_start: row.q_end]]
        if len(t) == 0 or len(q) == 0:
            continue
        if row.strand != 1:
            q = _COMPLEMENT[q[::-1]]
        recs.append(dict(query_name=qn, query=q.tobytes(), query_total_length=len(qseq[qn]), query_offset=row.q_start, query_is_rev=row.strand != 1,
                         target_name=tn, target=t.tobytes(), target_total_length=len(tseq[tn]), target_offset=row.r_start,
                         mashmap_estimated_identity=row.mashmap_estimated_identity, chain_id=row.chain_id, chain_length=row.chain_length,
                         chain_pos=row.chain_pos, mapping_query_span=row.q_end - row.q_start, row_index=row_index))
    return recs


def _phase_line(line: bytes) -> bytes:
    """Aligner::processMappingRecord (computeAlignments.hpp:486-516) re-emits a line that has a cg:Z: field as its whitespace-
    separated fields joined by single tabs: the trailing tab do_biwfa_alignment writes before the newline disappears. Lines
    without a CIGAR field (SAM records) pass through unchanged."""
    if not line:
        return line
    f = line.split()
    return b"\t".join(f) + b"\n" if any(x.startswith(b"cg:Z:") for x in f) else line


def align(paf: bytes, targets, queries, params: Params = None, device: int = 0, aligner=None, batch_records: int = 4096, per_row: bool = False):
    """Alignment phase over mapping PAF text -> (alignment PAF text, stats). per_row: return one bytes object per non-empty input
    row instead (b"" for rows that were skipped or filtered), for callers that scatter rows over GPUs."""
    P = (params or Params()).resolved()
    recs = records_from_paf(paf, targets, queries, P)
    own = aligner is None
    if own:
        aligner = wb.Aligner(device)
    lines, status = [], []
    for b in range(0, len(recs), batch_records):
        ln, st = aligner.biwfa_paf_batch(recs[b: b + batch_records], min_identity=P.min_identity, min_alignment_length=P.min_alignment_length,
                                         min_block_identity=P.min_block_identity, disable_chain_patching=P.disable_chain_patching,
                                         term_group=P.term_group)
        lines += [_phase_line(x) for x in ln]
        status += st
    if own:
        aligner.close()
    # the reference's processed_alignment_length (computeAlignments.hpp:480,528): qEndPos - qStartPos of every processed record
    aligned_bp = sum(r["mapping_query_span"] for r in recs)
    if per_row:
        rows = [b""] * sum(1 for ln in paf.split(b"\n") if ln)
        for r, ln in zip(recs, lines):
            rows[r["row_index"]] = ln
        lines = rows
    return (lines if per_row else b"".join(lines)), {"records": len(recs), "written": sum(1 for x in lines if x), "aligned_bp": aligned_bp, "status": status}


def wfmash(targets, queries, params: Params = None, device: int = 0):
    """map() then align(): the default `wfmash target.fa query.fa` run over in-memory sequences -> (alignment PAF, stats)."""
    P = (params or Params()).resolved()
    m = map(targets, queries, P, device)
    paf, st = align(m.paf, targets, queries, P, device)
    st["map"] = m.stats
    st["mapping_paf"] = m.paf
    return paf, st
