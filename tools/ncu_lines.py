"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv | python tools/ncu_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
fname, hdr, data = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ie = r.index("Instructions Executed"); ns = r.index("# Samples"); continue
    if hdr is None or r[0] in ("", "Function Name"): continue
    try: data.append((int(float(r[ie])), int(float(r[ns])), fname, r[0], r[1].strip()[:100]))
    except ValueError: pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print("total warp-instructions", tot, "samples", tots)
for d in sorted(data, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100 * d[0] / tot, 100 * d[1] / max(tots, 1), d[2], d[3], d[4]))
