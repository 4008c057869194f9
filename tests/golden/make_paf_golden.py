#!/usr/bin/env python3
"""Regenerates tests/golden/paf_do_biwfa.json.gz: the PAF text the UNMODIFIED reference's
wflign::wavefront::do_biwfa_alignment (src/common/wflign/src/wflign.cpp:108-483, compiled in place into
oracle/_ref/libwflignref.so with -march=x86-64-v3, i.e. its AVX2 extend kernels = term_group 8) writes for the
records of tests.util.paf_records(), once per filter set of tests.util.PAF_FILTER_SETS.
Run in the build container only (needs oracle/_ref, which needs /root/reference)."""
import gzip, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import util

R = util.load_wflign_ref()
assert R is not None, "build oracle/_ref first (make -C oracle)"
recs = util.paf_records()
out = []
for kw in util.PAF_FILTER_SETS:
    out.append([util.ref_paf(R, r, **kw).decode() for r in recs])
with gzip.GzipFile(os.path.join(HERE, "paf_do_biwfa.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"term_group": 8, "penalties": list(util.WFMASH_PEN), "filter_sets": util.PAF_FILTER_SETS, "lines": out}).encode())
print("records", len(recs), "lines written", [sum(1 for l in s if l) for s in out])

# the SAM branch (write_alignment_sam + MD tag) of the same unmodified function; lines kept as SHA-256 + first 9 columns
import hashlib
sam = []
for kw in util.SAM_SETS:
    rows = []
    for r in recs:
        ln = util.ref_sam(R, r, **kw)
        rows.append({"head": b"\t".join(ln.split(b"\t")[:5]).decode(), "sha": hashlib.sha256(ln).hexdigest(), "tail": ln[-60:].decode()})
    sam.append(rows)
with gzip.GzipFile(os.path.join(HERE, "sam_do_biwfa.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"term_group": 8, "sets": util.SAM_SETS, "lines": sam}).encode())
print("SAM records written", [sum(1 for l in s_ if l["head"]) for s_ in sam])
