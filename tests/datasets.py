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


def load(key):
    return read_fasta(path(key))


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
