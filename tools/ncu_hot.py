"""summarise an ncu --page source --print-source cuda,sass --csv export: hottest CUDA source lines by stall samples"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ''
lines = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 8 and r[0] not in ('', 'Line No'):
        try:
            si = hdr.index('# Samples')
            lines.append((int(r[si]), cur_file, r[0], r[1], int(r[hdr.index('Instructions Executed')])))
        except Exception:
            pass
tot = sum(l[0] for l in lines)
print('total samples', tot)
for s, f, ln, src, ie in sorted(lines, reverse=True)[:top]:
    print(f"{s:7d} {100*s/tot:5.1f}%  inst {ie:10d}  {f}:{ln:>4}  {src.strip()[:110]}")
