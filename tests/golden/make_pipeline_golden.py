#!/usr/bin/env python3
"""Regenerates tests/golden/pipeline_reference.json.gz: for every case of tests.pipeutil.PIPELINE_CASES the mapping PAF and
the alignment PAF that the reference-side pieces produce (tests/pipeutil.py::expected: oracle mapping restatement ->
UNMODIFIED mappingFilter.hpp / mappingOutput.hpp (libfilterref.so) -> UNMODIFIED do_biwfa_alignment (libwflignref.so,
AVX2 build = term_group 8)). The mapping PAF is stored in clear, every alignment line as its first 12 columns + SHA-256.
Run in the build container only (needs oracle/_ref, which needs /root/reference)."""
import gzip, hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import util, pipeutil

fref, wref = util.load_ref("libfilterref.so"), util.load_wflign_ref()
assert fref is not None and wref is not None, "build oracle/_ref first (make -C oracle)"
cases = []
for name, gen, prm in pipeutil.PIPELINE_CASES:
    seqs = pipeutil.case(**gen)
    mp, lines = pipeutil.expected(seqs, pipeutil.params(prm), util.load_oracle(), fref, wref)
    cases.append({"name": name, "mapping_paf": mp.decode(), "lines": [pipeutil.line_digest(l) for l in lines]})
    print(name, len(mp.splitlines()), "mappings,", sum(1 for l in lines if l), "alignment lines")
with gzip.GzipFile(os.path.join(HERE, "pipeline_reference.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"cases": cases}).encode())
