"SASS opcode histogram (executed instructions + stall samples) from `ncu --page source --csv` output."
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
src, ie, smp = ci['Source'], ci['Instructions Executed'], ci['# Samples']
ops, samp = collections.Counter(), collections.Counter()
tot = ts = 0
for r in rows[2:]:
    try:
        n = int(r[ie]); s = int(r[smp])
    except Exception:
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[src])
    op = m.group(2) if m else '?'
    ops[op] += n; samp[op] += s; tot += n; ts += s
print('total warp-instructions', tot, 'samples', ts)
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 28):
    print(f'{100*v/tot:5.1f}% inst  {100*samp[k]/max(ts,1):5.1f}% samples  {k}')
