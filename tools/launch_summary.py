"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the last full iteration."""
import csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def ms(row):
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    return v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
names = [(re.sub(r'\(.*', '', row['Kernel Name'])[:70], ms(row)) for row in rows]
idx = [i for i, n in enumerate(names) if 'post_kernel' in n[0]]
last = names[idx[-2] + 1: idx[-1] + 1]
tot = sum(v for _, v in last)
for n, v in last:
    print("%-72s %9.3f ms %5.1f%%" % (n, v, 100 * v / tot))
print("total %.3f ms over %d launches" % (tot, len(last)))
