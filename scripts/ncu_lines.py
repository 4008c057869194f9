"""Per-source-line share of the warp-stall samples of one kernel of an ncu report (needs -lineinfo + --import-source on):
   python scripts/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; fname = ""; acc = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= si or r[0] == "": continue
    try: acc.append((float(r[si]), float(r[ie]), fname, r[0], r[1].strip()[:120]))
    except ValueError: pass
tot = sum(a[0] for a in acc) or 1; toti = sum(a[1] for a in acc) or 1
print(f"kernel {kern}: {int(tot)} samples, {int(toti)} warp instructions")
for s, i, f, ln, src in sorted(acc, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% inst  {f}:{ln:>4s}  {src}")
