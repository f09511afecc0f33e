"""Per-source-line hot spots from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K`.
Usage: python tools/ncu_source_hotspots.py <mix.csv> [top]  -> table of (samples, share, instructions, file:line, source)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, cur_file, hdr = [], "", None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 6 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":     # a source line (SASS rows carry an address)
        try:
            s, ie = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        out.append((s, ie, f"{cur_file}:{r[0]}", r[1].strip()[:110]))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print(f"total samples {tot}, warp instructions {toti}, source lines {len(out)}")
for s, ie, loc, src in sorted(out, reverse=True)[:top]:
    print(f"{s:7d} {100*s/tot:5.1f}%  inst {100*ie/toti:5.1f}%  {loc:22s} {src}")
