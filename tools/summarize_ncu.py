"""Turns raw ncu output (gpurun_out/) into the small CSV summaries committed under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.csv> "<title>"
  python tools/summarize_ncu.py metrics  <report.ncu-rep> <out.csv> "<title>"
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
    'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
    'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'sm__cycles_elapsed.max',
]


def launches(src, dst, title):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith('==')) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] in ('ns', 'nsecond') else v * 1e3 if r[ui] in ('ms', 'msecond') else v
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# {title}\n# source: ncu --metrics gpu__time_duration.sum --clock-control none (cold cache, serialised launches)\n')
        f.write(f'# total {total:.1f} us over {n} launches\nlaunches,total_us,avg_us,share_pct,kernel\n')
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{c},{t:.1f},{t / c:.2f},{100 * t / total:.1f},"{k}"\n')
    print(open(dst).read())


def metrics(src, dst, title):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, 'w') as f:
        f.write(f'# {title}\n# source: ncu --set full --clock-control none; one column per profiled launch\n')
        f.write('metric,unit,' + ','.join(f'launch{i}' for i in range(len(data))) + '\n')
        ki = hdr.index('Kernel Name')
        f.write('Kernel Name,,' + ','.join('"' + d[ki] + '"' for d in data) + '\n')
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f'{m},{units[i]},' + ','.join('"' + d[i] + '"' for d in data) + '\n')
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'metrics': metrics}[sys.argv[1]](*sys.argv[2:5])
