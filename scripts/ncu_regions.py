"Stall samples of one kernel in an .ncu-rep, aggregated over buckets of SASS instructions (source page)."
import csv, collections, subprocess, sys
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 150
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
allagg = collections.Counter()
for b0 in range(0, len(data), B):
    chunk = data[b0:b0 + B]
    s = sum(int(r[ix['# Samples']]) for r in chunk)
    ex = sum(int(r[ix['Instructions Executed']]) for r in chunk)
    agg = collections.Counter()
    for r in chunk:
        for h in stalls: agg[h] += int(r[ix[h]])
    allagg.update(agg)
    def op(r):
        t = r[ix['Source']].split()
        return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops = collections.Counter(op(r) for r in chunk)
    notable = [k for k in ('LDTM', 'UTCHMMA', 'UTMALDG', 'MUFU', 'SYNCS', 'STS', 'LDS', 'STG', 'LDG', 'BAR', 'FMNMX', 'FADD', 'FSEL') if ops.get(k)]
    print(f"{b0:5d} samples {s:5d} ({100*s/tot:4.1f}%) exec {ex/1e6:6.2f}M  {[(k.replace('stall_',''),v) for k,v in agg.most_common(3)]}  {[(k,ops[k]) for k in notable]}")
print([(k.replace('stall_', ''), v) for k, v in allagg.most_common(12)])
