"""Top stall sites of an `ncu --page source --csv` export (SASS view): python scripts/ncu_src_top.py file.csv [kernel-substring] [n]"""
import csv, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ''; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'hdr': None, 'rows': []}; blocks.append(cur)
    elif cur is not None and r and r[0] == 'Address': cur['hdr'] = r
    elif cur is not None and cur['hdr'] and r: cur['rows'].append(r)
for b in blocks:
    if want not in b['name']: continue
    h = b['hdr']; ix = {k: i for i, k in enumerate(h)}
    samp = ix['# Samples']; src = ix['Source']; ex = ix['Instructions Executed']
    stall_cols = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
    tot = sum(int(r[samp] or 0) for r in b['rows'])
    print('==', b['name'][:90], 'samples', tot, 'instr', len(b['rows']))
    agg = {k: sum(int(r[ix[k]] or 0) for r in b['rows']) for k in stall_cols}
    print('  stall totals:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    top = sorted(b['rows'], key=lambda r: -int(r[samp] or 0))[:topn]
    for r in top:
        st = {k[6:]: int(r[ix[k]] or 0) for k in stall_cols if int(r[ix[k]] or 0)}
        print('  %6s %5s ex=%-7s %-60s %s' % (r[0][-5:], r[samp], r[ex], r[src][:60], st))
