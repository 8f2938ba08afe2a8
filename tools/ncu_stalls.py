"""Warp-stall samples of one .ncu-rep summed over all source lines, by reason.  usage: python tools/ncu_stalls.py report.ncu-rep"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
hdr, tot, n = None, collections.Counter(), 0
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try: s = int(r[hdr.index("# Samples")])
    except ValueError: continue
    n += s
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit(): tot[h[6:]] += int(r[i])
print("samples", n)
for k, v in tot.most_common(10): print("%-18s %6d %5.1f%%" % (k, v, 100 * v / max(n, 1)))
