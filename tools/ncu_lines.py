"""Top source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, data = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue   # source-level rows only
    try: s = int(r[hdr.index("# Samples")])
    except ValueError: continue
    stalls = {h[6:]: int(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit()}
    main = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    data.append((s, fname, r[0], r[1].strip()[:90], main, r[hdr.index("Instructions Executed")]))
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, f, ln, src, main, ie in sorted(data, reverse=True)[:top]:
    print("%5d %4.1f%% %s:%s [%s] inst=%s | %s" % (s, 100 * s / max(tot, 1), f, ln, ",".join("%s=%d" % m for m in main), ie, src))
