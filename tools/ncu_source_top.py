"""Digest an `ncu -i rep --page source --csv` export of ONE launch into a short text: warp-stall samples grouped by how
often each SASS instruction executed (= which warp role runs it: epilogue warps execute once per warp-tile, the MMA
issuer once per tile / tap, the producer once per CTA-tile) and the top stall sites.
usage: python tools/ncu_source_top.py source.csv [top_n]"""
import csv
import sys
from collections import defaultdict


def f(x):
  try:
    return float(x)
  except Exception:
    return 0.0


def main():
  path = sys.argv[1]
  top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
  rows = list(csv.reader(open(path)))
  hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
  h = rows[hi]
  idx = {c: i for i, c in enumerate(h)}
  data = [r for r in rows[hi + 1:] if len(r) == len(h)]
  # some exports carry the listing twice (SASS + a second view): keep the first copy
  addrs = [r[idx['Address']] for r in data]
  if len(addrs) > 1 and addrs[0] in addrs[1:]:
    data = data[:addrs.index(addrs[0], 1)]
  stall_cols = [c for c in h if c.startswith('stall_') and 'Not' not in c]
  total = sum(f(r[idx['# Samples']]) for r in data)
  print('# %s: %d SASS instructions, %d warp-stall samples' % (path.split('/')[-1], len(data), int(total)))
  groups = defaultdict(lambda: [0.0, 0])
  for r in data:
    g = groups[r[idx['Instructions Executed']]]
    g[0] += f(r[idx['# Samples']])
    g[1] += 1
  print('# samples by execution count of the instruction (warp role):')
  for k, (smp, n) in sorted(groups.items(), key=lambda kv: -kv[1][0])[:8]:
    print('#   executed %8s x : %6d samples (%4.1f %%) over %4d instructions' % (k, int(smp), 100.0 * smp / max(total, 1), n))
  print('# top stall sites: index, SASS, samples, executions, dominant stall reasons')
  for i, r in sorted(enumerate(data), key=lambda t: -f(t[1][idx['# Samples']]))[:top_n]:
    st = sorted(((c, f(r[idx[c]])) for c in stall_cols), key=lambda t: -t[1])[:2]
    st = ', '.join('%s %d' % (c.replace('stall_', ''), int(v)) for c, v in st if v > 0)
    print('%5d  %-64s %6d %8s  %s' % (i, r[idx['Source']].strip()[:64], int(f(r[idx['# Samples']])), r[idx['Instructions Executed']], st))


if __name__ == '__main__':
  main()
