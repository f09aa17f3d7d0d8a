"""Turn ncu outputs under gpurun_out/ into the small text/JSON summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r2.csv profiles/launches_r2_summary.md 32 [config]   (32 = bench batch)
      also writes profiles/launches_r2_traffic.json (dram bytes per conv_tc_kernel launch), stamped with the sha1 of the
      liboctseg.so it was measured on: bench.py only quotes it as roofline.traffic when the library is the same build
  python tools/summarize_ncu.py full gpurun_out/prof_x.ncu-rep profiles/prof_x_r1.txt
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']


def launches(src, dst, batch=None, config='ensemble'):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    ki, mi, vi, ui, ii = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('ID')
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        if r[mi].startswith('gpu__time'):
            v = v / 1e3 if u.startswith('n') else (v * 1e3 if u.startswith('m') else v)      # -> us
        else:
            v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        per[r[ii]][r[mi]] = v
        names[r[ii]] = r[ki].split('(')[0].replace('void ', '').split('<')[0]   # template specialisations aggregate
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for i, m in per.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += m.get('gpu__time_duration.sum', 0)
        a[2] += m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0)
    tot = sum(a[1] for a in agg.values())
    out = ['# ncu launch list summary (cold-cache, serialised launches: compare SHARES, not absolutes)', '',
           f'source: {src}; {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time', '',
           '| kernel | launches | time (ms) | share | DRAM bytes/launch (MB) |', '|---|---|---|---|---|']
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'| {k} | {a[0]} | {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f}% | {a[2] / max(a[0], 1) / 1e6:.2f} |')
    open(dst, 'w').write('\n'.join(out) + '\n')
    tc = next((v for k, v in agg.items() if k.split('::')[-1] == 'conv_tc_kernel'), None)
    if tc:
        import hashlib
        import os
        lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oct_segmentation_b200', 'liboctseg.so')
        sha = hashlib.sha1(open(lib, 'rb').read()).hexdigest()[:16] if os.path.exists(lib) else ''
        json.dump({'kernel': 'conv_tc_kernel', 'launches': tc[0], 'dram_bytes_per_launch': tc[2] / tc[0],
                   'share_of_kernel_time': tc[1] / tot, 'source': src, 'batch': int(batch) if batch else None, 'config': config,
                   'lib_sha16': sha}, open(dst.replace('_summary.md', '_traffic.json'), 'w'))
    print('\n'.join(out))


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f'# ncu --set full summary of {src}', '']
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        out.append(f"kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for h, v, u in zip(hdr, vals, units):
            if any(h == k or h.startswith(k + ' ') for k in KEYS):
                out.append(f'  {h} [{u}] = {v}')
        for h, v in zip(hdr, vals):
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(v or 0) > 0.2:
                out.append(f'  {h} = {v}')
        out.append('')
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:])
