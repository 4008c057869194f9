#!/usr/bin/env python3
"""Regenerates tests/golden/chain_reference.json.gz from the UNMODIFIED reference chain-merge code compiled in place
(oracle/_ref/libfilterref.so = src/map/include/mappingFilter.hpp mergeMappingsInRangeWithChains behind
oracle/ref_filter_driver.cpp): per case of tests/test_chain_cpu.py::CASES the SHA-256 of the reordered input mappings, of
the merged mappings and of their ChainInfo, the per-query offsets, and the first merged mappings in clear.
Run in the build container only (needs oracle/_ref, which needs /root/reference)."""
import gzip, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import util, chainutil
from tests.test_chain_cpu import CASES, digest, reference

ref = util.load_ref("libfilterref.so")
assert ref is not None, "build oracle/_ref first (make -C oracle)"
cases = []
for seed, gen, prm in CASES:
    m, off = chainutil.batch(seed, **gen)
    r_m, r_merged, r_info, r_mo = reference(ref, m, off, gen.get("w", 1000), **prm)
    cases.append({"seed": seed, "n_in": int(len(m)), "n_merged": int(len(r_merged)), "merged_offset": r_mo.tolist(), "sha_reordered": digest(r_m),
                  "sha_merged": digest(r_merged), "sha_chain_info": digest(r_info), "head": [[int(x) for x in r] for r in r_merged[:5].tolist()]})
    print(seed, len(m), "->", len(r_merged), "max chain", int(r_info["chainLen"].max()) if len(r_info) else 0)
with gzip.GzipFile(os.path.join(HERE, "chain_reference.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"cases": cases}).encode())
