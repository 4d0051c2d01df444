"""Join an ncu launch list (gpurun_out/launches.csv) with the op trace (step_calls.json)."""
import csv, json, re, collections, sys
lines = open('gpurun_out/launches.csv').readlines()
hi = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = [r for r in csv.DictReader(lines[hi:])]
mine = [r for r in rows if 'b200sr' in r['Kernel Name']]
calls = json.load(open('gpurun_out/step_calls.json'))
assert len(mine) == len(calls), (len(mine), len(calls))
def dur(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1000.0 if u.startswith('n') else v
agg = collections.OrderedDict()
for r, c in zip(mine, calls):
    k = re.sub(r'^(void )?b200sr::', '', r['Kernel Name'].split('(')[0])
    a = agg.setdefault((k, c), [0, 0.0]); a[0] += 1; a[1] += dur(r)
tot = sum(a[1] for a in agg.values())
print('total ms', tot / 1000)
byk = collections.Counter()
for (k, c), a in agg.items(): byk[k] += a[1]
for k, v in byk.most_common(): print(f'{k:36s} {v/1000:8.3f} ms {100*v/tot:5.1f}%')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for (k, c), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print(f'{k[:20]:20s} {c[:54]:54s} n={a[0]:4d} tot={a[1]/1000:7.3f} ms avg={a[1]/a[0]:7.1f} us')
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write('\n'.join(f'{k}\t{c}\t{a[0]}\t{a[1]:.1f}us' for (k, c), a in agg.items()))
