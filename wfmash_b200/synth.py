"""Seeded synthetic genomes / mapping records of the shapes BASELINE.json names (there is no network
and /root/reference does not exist on the GPU box, so bench.py and the full-size tests use these).

Model: i.i.d. uniform ACGT root; a derived sequence applies per-base independent
events at total rate d split substitution : insertion : deletion = 8 : 1 : 1, indel length
geometric with mean 3. numpy's PCG64 seeded generator replaces xoshiro256** (any fixed, seeded
stream serves; the exact generator is not part of the metric)."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_seq(n, rng):
    return _ACGT[rng.integers(0, 4, size=n)]


def mutate(seq, d, rng, sub_frac=0.8, ins_frac=0.1, indel_mean=3.0):
    """Return a mutated copy of seq (uint8 array) with total per-base event rate d."""
    n = int(seq.shape[0])
    if n == 0 or d <= 0:
        return seq.copy()
    r = rng.random(n)
    sub = r < d * sub_frac
    ins = (r >= d * sub_frac) & (r < d * (sub_frac + ins_frac))
    dele = (r >= d * (sub_frac + ins_frac)) & (r < d)
    out = seq.copy()
    # substitutions: always to a different base
    ns = int(sub.sum())
    if ns:
        code = np.searchsorted(_ACGT, out[sub])  # A,C,G,T sorted ascending in ASCII
        out[sub] = _ACGT[(code + rng.integers(1, 4, size=ns)) % 4]
    keep = np.ones(n, dtype=bool)
    dpos = np.flatnonzero(dele)
    if dpos.size:
        dl = rng.geometric(1.0 / indel_mean, size=dpos.size)
        diff = np.zeros(n + 1, dtype=np.int64)
        np.add.at(diff, dpos, 1)
        np.add.at(diff, np.minimum(dpos + dl, n), -1)
        keep = np.cumsum(diff[:n]) == 0
    ins_len = np.zeros(n, dtype=np.int64)
    ipos = np.flatnonzero(ins)
    if ipos.size:
        ins_len[ipos] = rng.geometric(1.0 / indel_mean, size=ipos.size)
    contrib = keep.astype(np.int64) + ins_len
    total = int(contrib.sum())
    res = np.empty(total, dtype=np.uint8)
    start = np.cumsum(contrib) - contrib
    kept_idx = np.flatnonzero(keep)
    res_is_kept = np.zeros(total, dtype=bool)
    res_is_kept[start[kept_idx]] = True
    res[start[kept_idx]] = out[kept_idx]
    n_ins = total - kept_idx.size
    if n_ins:
        res[~res_is_kept] = _ACGT[rng.integers(0, 4, size=n_ins)]
    return res


def mapping_records(n, seed, len_lo, len_hi, divergences, pad=0):
    """n synthetic mapping records as the aligner sees them (computeAlignments.hpp:195-303):
    (target slice = pattern, query slice = text), query length ~ U[len_lo, len_hi], the target is the
    query's source mutated at a divergence drawn from `divergences`, plus `pad` unrelated flanking bases
    on both target ends (wfmash pads the target by min(w,5000), parse_args.hpp:608)."""
    rng = np.random.default_rng(seed)
    recs = []
    for _ in range(n):
        qlen = int(rng.integers(len_lo, len_hi + 1))
        d = float(divergences[int(rng.integers(0, len(divergences)))])
        q = random_seq(qlen, rng)
        t = mutate(q, d, rng)
        if pad:
            t = np.concatenate([random_seq(pad, rng), t, random_seq(pad, rng)])
        recs.append((t.tobytes(), q.tobytes(), d))
    return recs


def genome(n_contigs, contig_len, seed):
    rng = np.random.default_rng(seed)
    return [random_seq(contig_len, rng) for _ in range(n_contigs)]


# ---- SURVEY 8(d)'s generator for the synthetic configs C4 / C5: xoshiro256** seed 42 -------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return x, z ^ (z >> 31)


def _xo_next(s):
    """One step of the scalar generator on a list of four Python ints (used for seeding and jump())."""
    r = (s[1] * 5) & _M64
    r = (((r << 7) | (r >> 57)) & _M64) * 9 & _M64
    t = (s[1] << 17) & _M64
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t
    s[3] = ((s[3] << 45) | (s[3] >> 19)) & _M64
    return r


_XO_JUMP = (0x180EC6D33CFD0ABA, 0xD5A61266F0C9392C, 0xA9582618E03FC9AA, 0x39ABDC4529B1661C)


def _xo_jump(s):
    """The generator's published jump(): 2^128 steps ahead, i.e. the next non-overlapping sub-stream."""
    acc = [0, 0, 0, 0]
    for word in _XO_JUMP:
        for b in range(64):
            if (word >> b) & 1:
                for i in range(4):
                    acc[i] ^= s[i]
            _xo_next(s)
    s[:] = acc


class Xoshiro256ss:
    """xoshiro256** (Blackman & Vigna), state seeded from `seed` through splitmix64 as its authors prescribe, run as `lanes` sub-streams:
    lane j is the seeded generator after j jump() calls (2^128 steps apart), and the output sequence interleaves the lanes
    (value i comes from lane i % lanes). numpy only vectorises the lanes; every number is a pure function of (seed, lanes), independent
    of the numpy version. Offers the three draws synth.mutate needs with numpy Generator's names."""

    def __init__(self, seed=42, lanes=256):
        x, s = seed & _M64, []
        for _ in range(4):
            x, z = _splitmix64(x)
            s.append(z)
        st = np.empty((4, lanes), dtype=np.uint64)
        for j in range(lanes):
            st[:, j] = s
            _xo_jump(s)
        self.s0, self.s1, self.s2, self.s3 = (st[i].copy() for i in range(4))
        self.lanes = lanes

    def raw(self, n):
        """n 64-bit outputs (whole rounds of `lanes` are consumed)."""
        L = self.lanes
        steps = (int(n) + L - 1) // L
        out = np.empty((steps, L), dtype=np.uint64)
        s0, s1, s2, s3 = self.s0, self.s1, self.s2, self.s3
        c5, c9 = np.uint64(5), np.uint64(9)
        for i in range(steps):
            r = s1 * c5
            r = ((r << np.uint64(7)) | (r >> np.uint64(57))) * c9
            out[i] = r
            t = s1 << np.uint64(17)
            s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t
            s3[:] = (s3 << np.uint64(45)) | (s3 >> np.uint64(19))
        return out.reshape(-1)[: int(n)]

    def random(self, n):
        """doubles in [0, 1): the top 53 bits."""
        return (self.raw(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def integers(self, lo, hi, size):
        """integers in [lo, hi): multiply-shift of the top 32 bits (ranges here are 3 or 4 wide)."""
        x = self.raw(size) >> np.uint64(32)
        return (lo + ((x * np.uint64(hi - lo)) >> np.uint64(32)).astype(np.int64))

    def geometric(self, p, size):
        """P(X = j) = (1 - p)^(j-1) p, j >= 1, by inversion against a table of exactly rounded products (no libm)."""
        q, th = 1.0, []
        for _ in range(200):
            q *= (1.0 - p)
            th.append(q)          # th[j-1] = P(X > j)
        asc = np.array(th[::-1])
        u = self.random(size)
        # X = 1 + #{j >= 1 : u < P(X > j)}
        return 1 + (len(th) - np.searchsorted(asc, u, side="right")).astype(np.int64)


def pansn_pangenome(shape, seed=42):
    """[(PanSN name, bytes)] of SURVEY 8(d)'s synthetic configs (scaled by the caller through the lengths):
      shape = dict(kind="C4", contigs=20, contig_len=50_000_000, ani=0.90): two genomes gA / gB of `contigs` contigs
              (gA#1#chr01 ...), each derived independently from one i.i.d. root with d = 1 - sqrt(ani);
      shape = dict(kind="C5", haplotypes=100, contig_len=50_000_000, ani=0.80): `haplotypes` haplotypes of one contig (hNNN#1#chr1)."""
    rng = Xoshiro256ss(seed)
    d = 1.0 - float(np.sqrt(shape["ani"]))
    out = []
    if shape["kind"] == "C4":
        roots = [random_seq(shape["contig_len"], rng) for _ in range(shape["contigs"])]
        for g in ("gA", "gB"):
            for c, root in enumerate(roots):
                out.append((f"{g}#1#chr{c + 1:02d}", mutate(root, d, rng).tobytes()))
    elif shape["kind"] == "C5":
        root = random_seq(shape["contig_len"], rng)
        for h in range(shape["haplotypes"]):
            out.append((f"h{h + 1:03d}#1#chr1", mutate(root, d, rng).tobytes()))
    else:
        raise ValueError(shape["kind"])
    return out
