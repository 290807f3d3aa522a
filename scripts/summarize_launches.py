"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
ui = hdr.index('Metric Unit')
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    if r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[ui], 1e-3)
    name = re.sub(r'\(.*', '', r[ki])[:90]
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print(f'total {total/1e3:.3f} ms over {sum(cnt.values())} launches')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{v/1e3:9.3f} ms {100*v/total:5.1f}%  x{cnt[k]:<4d} {k}')
