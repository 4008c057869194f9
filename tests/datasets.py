"""The reference's own example inputs (BASELINE.json configs C1-C3), shipped under tests/data/ as byte copies of
/root/reference/data/*.fa.gz (+ .fai / .gzi) so that they exist on the GPU box, and a tiny loader for them.

TEST / BENCH INFRASTRUCTURE ONLY: the product library takes sequences that are already in host memory (the reference reads its
FASTA through htslib/faidx on the host, src/common/faigz.h; FASTA / BGZF I/O is out of scope by SURVEY 8). BGZF is multi-member
gzip, which `gzip.open` reads."""
import gzip
import os

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

FILES = {
    "reads255": "reads.255bps.fa.gz",      # C1 queries: 8 reads of 255-257 bp
    "reference": "reference.fa.gz",        # C1 target: 1 sequence, 1 399 930 bp
    "lpa": "LPA.subset.fa.gz",             # C2: 8 sequences, 2 317 910 bp
    "yeast": "scerevisiae8.fa.gz",         # C3: 136 sequences, 96 255 507 bp, 8 PanSN groups
}


def path(key):
    return os.path.join(DATA, FILES[key])


def read_fasta(p):
    """[(name, bytes)] in file order; name = the header up to the first white space (faidx's rule), sequence bytes as stored
    (case kept: the library upper-cases / N-masks exactly where the reference does)."""
    out, name, parts = [], None, []
    with gzip.open(p, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    out.append((name, b"".join(parts)))
                name, parts = line[1:].split()[0].decode(), []
            else:
                parts.append(line.rstrip(b"\r\n"))
    if name is not None:
        out.append((name, b"".join(parts)))
    return out


# SURVEY 8(d)'s synthetic configs C4 / C5 at a scale the CPU reference finishes in minutes here (xoshiro256** seed 42, PanSN names,
# substitution : insertion : deletion = 8 : 1 : 1; wfmash_b200.synth.pansn_pangenome). Generated, not shipped: every byte is a pure
# function of the shape, so the build container (fixture) and the GPU box (test / bench) see the same sequences.
SYNTH = {
    "synth_c4s": dict(kind="C4", contigs=8, contig_len=2_500_000, ani=0.90),     # 2 genomes x 8 contigs x 2.5 Mbp (C4 is 2 x 20 x 50 Mbp)
    "synth_c5s": dict(kind="C5", haplotypes=5, contig_len=1_000_000, ani=0.80),  # 5 haplotypes x 1 Mbp (C5 is 100 x 50 Mbp)
}


def load(key):
    if key in SYNTH:
        return synthetic(key)
    return read_fasta(path(key))


# sha256 over name + sequence of what the generator produced when the reference fixture was made: a different numpy / platform that
# changed a single base would otherwise show up as an unexplained parity failure
SYNTH_SHA = {
    "synth_c4s": "f4b9eaddfe05d2a05110099df1223bedd31004f7c2005285d9a89c2029bdc9c3",
    "synth_c5s": "17c50ab4113f6955974881aec105672a9c7f4638871674bcd421e536e40428c9",
}


def synthetic(key):
    import hashlib
    from wfmash_b200 import synth
    seqs = synth.pansn_pangenome(SYNTH[key], seed=42)
    got = hashlib.sha256(b"".join(n.encode() + b"\n" + x + b"\n" for n, x in seqs)).hexdigest()
    assert got == SYNTH_SHA[key], f"{key}: the generated sequences are not the ones the reference fixture was made from ({got})"
    return seqs


def yeast_subset(seqs, genomes=2, chroms=None):
    """The sequences of the first `genomes` PanSN groups (in file order) of scerevisiae8, optionally only chromosomes whose
    last PanSN field is one of `chroms`."""
    keep, out = [], []
    for n, s in seqs:
        g = n[: n.rfind("#")]
        if g not in keep:
            if len(keep) == genomes:
                continue
            keep.append(g)
        if chroms is None or n[n.rfind("#") + 1:] in chroms:
            out.append((n, s))
    return out
