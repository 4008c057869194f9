#!/usr/bin/env python3
"""Regenerates tests/golden/wfa_utest.{seq,biwfa.affine2p.alg,affine2p.alg} from the reference's own
known-answer vectors (deps/WFA2-lib/tests/wfa.utest.seq and tests/wfa.utest.check/*.alg, produced by
the reference's align_benchmark with penalties 0,4,6,2,24,1; see tests/wfa.utest.sh:30-50).
These are DATA fixtures (sequence pairs + expected score/CIGAR), not reference source code.
Also writes wfa_wfmash_pen.tsv: CIGARs for the same pairs under wfmash's penalties (0,5,8,2,24,1,
wflign.cpp:136-144), produced by running the unmodified reference (oracle/_ref/libwfa2ref.so).
Run in the build container only (needs /root/reference)."""
import ctypes, gzip, os, shutil, sys
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/deps/WFA2-lib/tests"
shutil.copy(f"{REF}/wfa.utest.seq", f"{HERE}/wfa_utest.seq")
shutil.copy(f"{REF}/wfa.utest.check/test.biwfa.affine2p.alg", f"{HERE}/wfa_utest.biwfa.affine2p.alg")
shutil.copy(f"{REF}/wfa.utest.check/test.affine2p.alg", f"{HERE}/wfa_utest.affine2p.alg")
lib = ctypes.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libwfa2ref.so"))
seqs = [l.rstrip("\n") for l in open(f"{HERE}/wfa_utest.seq")]
with open(f"{HERE}/wfa_wfmash_pen.tsv", "w") as out:
    for i in range(0, len(seqs), 2):
        p, t = seqs[i][1:].encode(), seqs[i + 1][1:].encode()
        row = []
        for mode in (3, 0):  # ultralow (biWFA), high (unidirectional)
            buf = ctypes.create_string_buffer(2 * (len(p) + len(t)) + 16)
            n, sc = ctypes.c_int(), ctypes.c_int()
            st = lib.ref_wfa_end2end(p, len(p), t, len(t), 5, 8, 2, 24, 1, mode, buf, len(buf),
                                     ctypes.byref(n), ctypes.byref(sc))
            assert st == 0
            row.append(buf.raw[: n.value].decode())
        out.write("\t".join(row) + "\n")
for f in ("wfa_utest.seq", "wfa_utest.biwfa.affine2p.alg", "wfa_utest.affine2p.alg", "wfa_wfmash_pen.tsv"):
    with open(f"{HERE}/{f}", "rb") as fi, gzip.GzipFile(f"{HERE}/{f}.gz", "wb", mtime=0) as fo:
        fo.write(fi.read())
    os.remove(f"{HERE}/{f}")
