"""Stall samples per CUDA source line of a kernel in an ncu report: python scratch/ncu_lines.py rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
per = collections.Counter(); src = {}; lsb = collections.Counter()
hdr = None
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try:
        n = int(r[hdr.index('# Samples')])
    except ValueError:
        continue
    line = r[0]
    if r[1].strip(): src[line] = r[1].strip()[:100]
    per[line] += n
    try: lsb[line] += int(r[hdr.index('stall_long_sb')])
    except ValueError: pass
tot = sum(per.values())
print('total samples', tot)
for line, n in per.most_common(top):
    print(f'{n:7d} {100*n/tot:5.1f}%  long_sb {lsb[line]:6d}  L{line}: {src.get(line, "")}')
