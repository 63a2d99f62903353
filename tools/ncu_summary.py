"""Digest an `ncu --page raw --csv` export: one line per launch (kernel + template args, duration, tensor-pipe %, DRAM
bytes / GB/s / % of peak, L2 / L1 throughput %), plus time-weighted figures per kernel family.
usage: python tools/ncu_summary.py raw.csv [--json out.json] [--peak-gbs 6550]"""
import csv
import json
import re
import sys
from collections import OrderedDict


def num(x):
  try:
    return float(str(x).replace(',', ''))
  except Exception:
    return float('nan')


def col(row, *frags):
  for k, v in row.items():
    if all(f in k for f in frags):
      return num(v)
  return float('nan')


def main():
  path = sys.argv[1]
  peak = 6550.1
  out_json = None
  for i, a in enumerate(sys.argv):
    if a == '--peak-gbs':
      peak = float(sys.argv[i + 1])
    if a == '--json':
      out_json = sys.argv[i + 1]
  rows = list(csv.DictReader(open(path)))
  rows = [r for r in rows if r.get('Kernel Name') and not r['ID'].strip() == '']      # 2nd row holds units
  fam = OrderedDict()
  lines = []
  for r in rows:
    name = r['Kernel Name']
    if not name or name == '':
      continue
    m = re.match(r'(?:void )?(?:immb::)?([A-Za-z0-9_]+)(<[^>]*>)?', name)
    short = (m.group(1) + (m.group(2) or '')) if m else name[:40]
    dur_ns = col(r, 'gpu__time_duration.sum')
    if dur_ns != dur_ns:
      continue
    tensor = col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
    if tensor != tensor:
      tensor = col(r, 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed')
    rd, wr = col(r, 'dram__bytes_read.sum'), col(r, 'dram__bytes_write.sum')
    # ncu reports bytes with a unit row; raw page gives plain numbers in the unit of the 2nd header row (bytes or Mbyte):
    l1 = col(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed')
    lts = col(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')
    dram_pct = col(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed')
    regs = col(r, 'launch__registers_per_thread')
    lines.append((short, dur_ns, tensor, rd, wr, dram_pct, lts, l1, regs, r.get('Grid Size', '')))
    f = fam.setdefault(short, {'n': 0, 'ns': 0.0, 'tensor_w': 0.0, 'rd': 0.0, 'wr': 0.0, 'dram_w': 0.0})
    f['n'] += 1
    f['ns'] += dur_ns
    f['tensor_w'] += (tensor if tensor == tensor else 0.0) * dur_ns
    f['rd'] += rd if rd == rd else 0.0
    f['wr'] += wr if wr == wr else 0.0
    f['dram_w'] += (dram_pct if dram_pct == dram_pct else 0.0) * dur_ns
  units = None
  with open(path) as fh:
    rd_ = csv.reader(fh)
    hdr = next(rd_)
    u = next(rd_)
    units = dict(zip(hdr, u))
  def unit_of(frag):
    for k, v in units.items():
      if frag in k:
        return v
    return ''
  byte_unit = unit_of('dram__bytes_read.sum')
  mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(byte_unit, 1.0)
  tunit = unit_of('gpu__time_duration.sum')
  tmult = {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'nsecond': 1.0, 'usecond': 1e3, 'msecond': 1e6}.get(tunit, 1.0)
  print('# %d launches; bytes unit %r, time unit %r' % (len(lines), byte_unit, tunit))
  print('%-44s %9s %8s %9s %9s %7s %6s %6s %5s %s' % ('kernel', 'us', 'tensor%', 'rd MB', 'wr MB', 'GB/s', 'dram%', 'lts%', 'l1%', 'grid'))
  for short, ns, tensor, rd, wr, dram_pct, lts, l1, regs, grid in lines:
    ns *= tmult
    gbs = (rd + wr) * mult / ns if ns > 0 else 0.0
    print('%-44s %9.1f %8.1f %9.1f %9.1f %7.0f %6.1f %6.1f %5.1f %s' % (short[:44], ns / 1e3, tensor, rd * mult / 1e6, wr * mult / 1e6, gbs, dram_pct, lts, l1, grid))
  print('\n# per kernel (time-weighted)')
  total_ns = sum(f['ns'] for f in fam.values()) * tmult
  summary = OrderedDict()
  for k, f in sorted(fam.items(), key=lambda kv: -kv[1]['ns']):
    ns = f['ns'] * tmult
    gbs = (f['rd'] + f['wr']) * mult / ns if ns else 0.0
    summary[k] = {'launches': f['n'], 'us': ns / 1e3, 'share': ns / total_ns, 'tensor_pipe_pct': f['tensor_w'] / f['ns'] if f['ns'] else 0.0,
                  'dram_mb_per_launch': (f['rd'] + f['wr']) * mult / 1e6 / f['n'], 'dram_gbs': gbs, 'frac_of_hbm_peak': gbs / peak,
                  'dram_pct_of_peak_ncu': f['dram_w'] / f['ns'] if f['ns'] else 0.0}
    print('%-44s n=%3d %9.1f us %5.1f%%  tensor %5.1f%%  %8.1f MB/launch  %6.0f GB/s (%.2f of %.0f)' %
          (k[:44], f['n'], ns / 1e3, 100 * ns / total_ns, summary[k]['tensor_pipe_pct'], summary[k]['dram_mb_per_launch'], gbs, gbs / peak, peak))
  if out_json:
    json.dump(summary, open(out_json, 'w'), indent=1)


if __name__ == '__main__':
  main()
