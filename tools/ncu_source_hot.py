"""Hot source lines of one captured launch: python tools/ncu_source_hot.py <csv from `ncu -i rep --page source --csv --print-source cuda,sass ...`> [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        src = r[1]
        out.append((cur, int(r[0]), src[:120], int(d.get('Warp Stall Sampling (All Samples)') or 0), int(d.get('Instructions Executed') or 0)))
ts = sum(o[3] for o in out) or 1
ti = sum(o[4] for o in out) or 1
print('total stall samples', ts, ' warp instructions', ti)
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print('%-20s %4d  samples %5.1f%%  instr %5.1f%%  %s' % (o[0], o[1], 100 * o[3] / ts, 100 * o[4] / ti, o[2]))
