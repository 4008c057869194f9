#!/usr/bin/env python3
"""Regenerates tests/golden/filter_reference.json.gz from the UNMODIFIED reference filter code compiled in place
(oracle/_ref/libfilterref.so = src/map/include/mappingFilter.hpp + filter.hpp + mappingOutput.hpp behind
oracle/ref_filter_driver.cpp, which issues the calls of Map::filterSubsetMappings in the reference's order): per case of
tests/test_filter_cpu.py::CASES the SHA-256 of the surviving mappings, of the ChainInfo the reference pairs them with and
of the mapping PAF text, the per-query offsets, and the first PAF lines in clear.
Run in the build container only (needs oracle/_ref, which needs /root/reference)."""
import gzip, hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import util, chainutil
from tests.test_filter_cpu import CASES, digest, reference, paf_reference

ref = util.load_ref("libfilterref.so")
assert ref is not None, "build oracle/_ref first (make -C oracle)"
cases = []
for seed, gen, prm, groups in CASES:
    m, off = chainutil.batch(seed, **gen)
    r_out, r_info, r_oo = reference(ref, gen, prm, groups, m, off)
    txt = paf_reference(ref, gen, prm, r_out, r_info, r_oo)
    cases.append({"seed": seed, "n_in": int(len(m)), "n_out": int(len(r_out)), "out_offset": r_oo.tolist(), "sha_out": digest(r_out),
                  "sha_chain_info": digest(r_info), "sha_paf": hashlib.sha256(txt).hexdigest(), "paf_head": txt.decode().splitlines()[:3]})
    print(seed, len(m), "->", len(r_out), "mappings,", len(txt), "bytes of PAF")
with gzip.GzipFile(os.path.join(HERE, "filter_reference.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"cases": cases}).encode())
