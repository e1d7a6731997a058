"Aggregate an ncu launch list (gpu__time_duration) by kernel name for the last 1/Nth of the run."
import csv, collections, re, sys
path = sys.argv[1]; parts = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
n = len(rows) // parts
last = rows[(parts - 1) * n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for x in last:
    name = re.sub(r'\(.*', '', x['Kernel Name'])
    name = re.sub(r'void |dmg::|\(anonymous namespace\)::|<unnamed>::', '', name)
    v = float(x['Metric Value'].replace(',', '')) * (1e-3 if x['Metric Unit'] == 'ns' else 1)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v[1]/1e3:9.3f} ms {v[0]:5d}  {100*v[1]/tot:5.1f}%  {v[1]/v[0]:8.1f} us/launch  {k[:80]}')
print('total ms', tot / 1e3, 'launches', len(last))
