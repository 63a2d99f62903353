"""Build profiles/top_kernel.json (what bench.py's roofline block quotes from ncu) from the committed captures:
  pair kernel  : profiles/r2f_pair_52launches_ncu_full_raw.csv  (ncu --set full, 52 launches of one step)
  whole step   : profiles/r2g_step_metrics_raw.csv              (ncu --metrics ..., every launch of one step)
usage: python tools/make_top_kernel.py"""
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
  try:
    return float(str(x).replace(',', ''))
  except Exception:
    return float('nan')


def load(path):
  with open(path) as fh:
    rd = csv.reader(fh)
    hdr = next(rd)
    units = dict(zip(hdr, next(rd)))
    rows = [dict(zip(hdr, r)) for r in rd if len(r) == len(hdr)]
  bmult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units['dram__bytes_read.sum']]
  tmult = {'ns': 1.0, 'us': 1e3, 'ms': 1e6}.get(units['gpu__time_duration.sum'], 1.0)
  out = []
  for r in rows:
    m = re.match(r'(?:void )?(?:immb::)?([A-Za-z0-9_]+)(<[^>]*>)?', r['Kernel Name'])
    out.append({'name': m.group(1), 'targs': m.group(2) or '',
                'ns': num(r['gpu__time_duration.sum']) * tmult,
                'bytes': (num(r['dram__bytes_read.sum']) + num(r['dram__bytes_write.sum'])) * bmult,
                'tensor': num(r['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'])})
  return out


def build(hbm=None):
  """hbm: copy-bandwidth peak in GB/s (default: MEASURED_PEAKS.json of this pod)."""
  if hbm is None:
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    hbm = float(peaks.get('hbm_gbs', peaks.get('hbm_gbs_burst', 6550.0)))
  pair = [k for k in load(os.path.join(ROOT, 'profiles', 'r2f_pair_52launches_ncu_full_raw.csv')) if k['name'] == 'conv_tc2_pair_kernel']
  step = load(os.path.join(ROOT, 'profiles', 'r2g_step_metrics_raw.csv'))
  tot_ns = sum(k['ns'] for k in pair)
  fam = {}
  for k in pair:
    f = fam.setdefault(k['targs'], {'launches': 0, 'ns': 0.0, 'tw': 0.0})
    f['launches'] += 1
    f['ns'] += k['ns']
    f['tw'] += k['tensor'] * k['ns']
  hbm_fams = {}
  for k in step:
    if k['tensor'] > 1.0 or k['bytes'] < 8e6:          # tensor-core kernels and the tiny latency-bound ones are not HBM-bound
      continue
    f = hbm_fams.setdefault(k['name'] + k['targs'], {'launches': 0, 'ns': 0.0, 'bytes': 0.0})
    f['launches'] += 1
    f['ns'] += k['ns']
    f['bytes'] += k['bytes']
  step_ns = sum(k['ns'] for k in step)
  out = {
      'kernel': 'conv_tc2_pair_kernel<BN2,PASSES,BNR,F16> (persistent CTA-pair halo conv, tcgen05 cta_group::2, M=256): the 52 '
                'launches of one training step (forward + dgrad of every stride-1 3x3 layer and the 7x7 first layers), batch 64',
      'source': 'profiles/r2f_pair_52launches_ncu_full_raw.csv (ncu --set full --clock-control none --import-source on -k '
                'regex:conv_tc2_pair_kernel -c 52 python tools/profile_step.py 64); HBM kernels: profiles/r2g_step_metrics_raw.csv',
      'launches': len(pair),
      'dram_bytes_per_launch': sum(k['bytes'] for k in pair) / len(pair),
      'time_us_per_launch': tot_ns / len(pair) / 1e3,
      'tensor_pipe_pct_time_weighted': sum(k['tensor'] * k['ns'] for k in pair) / tot_ns,
      'tensor_pipe_by_family': {t: {'launches': f['launches'], 'us': round(f['ns'] / 1e3, 1), 'tensor_pipe_pct': round(f['tw'] / f['ns'], 1)}
                                for t, f in sorted(fam.items(), key=lambda kv: -kv[1]['ns'])},
      'hbm_peak_gbs': hbm,
      'hbm_kernels': {n: {'launches': f['launches'], 'us': round(f['ns'] / 1e3, 1), 'dram_mb_per_launch': round(f['bytes'] / f['launches'] / 1e6, 1),
                          'gbs': round(f['bytes'] / f['ns'], 0), 'frac_of_hbm_peak': round(f['bytes'] / f['ns'] / hbm, 3),
                          'share_of_step': round(f['ns'] / step_ns, 4)}
                      for n, f in sorted(hbm_fams.items(), key=lambda kv: -kv[1]['ns'])},
      'note': 'ncu launches are cold-cache and serialised (12.7 ms for the step vs 11.4 ms live): shares, not absolutes, carry over',
  }
  return out


def main():
  out = build()
  json.dump(out, open(os.path.join(ROOT, 'profiles', 'top_kernel.json'), 'w'), indent=1)
  print(json.dumps(out, indent=1)[:3000])


if __name__ == '__main__':
  main()
