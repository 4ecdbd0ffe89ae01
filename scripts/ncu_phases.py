"""Per-phase (barrier-delimited) summary of an ncu source-page export:  ncu -i X.ncu-rep --page source --csv > src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
px = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
iS = hdr.index('Source'); iN = hdr.index('# Samples'); iE = hdr.index('Instructions Executed')
seg = 0; agg = {}; tot_s = tot_e = 0
for r in rows[2:]:
    if len(r) < len(hdr) or r[iN] == '# Samples': continue
    s = r[iS].strip(); n = int(r[iN] or 0); e = int(r[iE] or 0)
    a = agg.setdefault(seg, [0, 0, 0, []])
    a[0] += n; a[1] += e; a[2] += 1; a[3].append((n, e, s))
    tot_s += n; tot_e += e
    if 'BAR.SYNC' in s or ('RET.' in s): seg += 1
print('total samples', tot_s, 'warp instr', tot_e, 'thread-instr/px', tot_e * 32 / px)
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
for k, v in agg.items():
    print(f'seg {k}: samples {v[0]} ({100*v[0]/max(tot_s,1):.1f}%) warp-instr {v[1]} ({v[1]*32/px:.2f} thr-instr/px) static {v[2]}')
    for n, e, s in sorted(v[3], reverse=True)[:top_n]: print('      ', n, e, s[:100])
