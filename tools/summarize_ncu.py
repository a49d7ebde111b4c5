"""Turn the ncu CSV logs of one round into the committed summaries under profiles/."""
import collections
import csv
import re
import sys


def read(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    for r in rd:
        rows.append(dict(zip(hdr, r)))
    return rows


def short(name):
    m = re.search(r'(\w+_kernel(?:<[^>]*>)?|vectorized_elementwise_kernel|\w+Kernel\w*|ncclDevKernel\w*)', name)
    s = m.group(1) if m else name[:60]
    return s.replace('(anonymous namespace)::', '')


def launches(path, out, steps=2):
    rows = read(path)
    # ncu --metrics prints one row per (launch, metric): keep gpu__time_duration
    per = [(int(r['ID']), short(r['Kernel Name']), float(r['Metric Value'].replace(',', ''))) for r in rows if r['Metric Name'] == 'gpu__time_duration.sum']
    unit = [r['Metric Unit'] for r in rows if r['Metric Name'] == 'gpu__time_duration.sum'][0]
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1.0)
    # the last step = the launches after the previous step's adamw_kernel up to and including the last adamw_kernel (the first
    # step also holds one-off work: optimizer-state zero fills, operand casts, allocator warm-up)
    ends = [i for i, p in enumerate(per) if p[1].startswith('adamw_kernel')]
    second = per[ends[-2] + 1:ends[-1] + 1] if len(ends) >= 2 else per[-(len(per) // steps):]
    agg = collections.OrderedDict()
    for _, k, v in second:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write('# ncu launch list summary (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n')
        f.write('# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv python tools/ncu_step.py sr_tiny 256 2\n')
        f.write('# last of %d steps: %d launches, %.2f ms of kernel time\n\n' % (steps, len(second), tot / 1e3))
        f.write('| kernel | launches | total us | share |\n|---|---:|---:|---:|\n')
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.1f%% |\n' % (k, n, us, 100 * us / tot))
    return agg, tot


def gemm_dram(path, out, steps=2):
    rows = read(path)
    by_id = collections.OrderedDict()
    for r in rows:
        d = by_id.setdefault(int(r['ID']), {'name': short(r['Kernel Name'])})
        d[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
    ids = list(by_id)
    ids = ids[-(len(ids) // steps):]          # GEMM launches of the last step (both steps launch the same number)

    def val(d, k, want):
        v, u = d[k]
        f = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'nsecond': 1e-9, 'usecond': 1e-6,
             'msecond': 1e-3, '%': 1}.get(u, 1)
        return v * f
    agg = collections.OrderedDict()
    for i in ids:
        d = by_id[i]
        a = agg.setdefault(d['name'], [0, 0.0, 0.0, 0.0, 0.0])
        t = val(d, 'gpu__time_duration.sum', 's')
        a[0] += 1
        a[1] += t
        a[2] += val(d, 'dram__bytes_read.sum', 'B') + val(d, 'dram__bytes_write.sum', 'B')
        a[3] += d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][0] * t
        a[4] += d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0] * t
    with open(out, 'w') as f:
        f.write('# tcgen05 GEMM launches of one train step (sr_tiny, B=256): DRAM traffic and pipe utilisation from ncu\n')
        f.write('# command: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,gpu__dram_throughput... -k regex:gemm_tc -c 600\n\n')
        f.write('| kernel (EPI: 0 store, 1 gelu, 2 residual, 3 gelu-grad, 4 atomic) | launches | time ms | DRAM GB | DRAM GB/s | tensor pipe active % | dram throughput % |\n|---|---:|---:|---:|---:|---:|---:|\n')
        T = B = 0.0
        for k, (n, t, b, tp, dp) in agg.items():
            f.write('| `%s` | %d | %.3f | %.3f | %.0f | %.1f | %.1f |\n' % (k, n, t * 1e3, b / 1e9, b / t / 1e9, tp / t, dp / t))
            T += t
            B += b
        f.write('| **all** | %d | %.3f | %.3f | %.0f | | |\n' % (sum(v[0] for v in agg.values()), T * 1e3, B / 1e9, B / T / 1e9))
    per_launch = B / max(1, sum(v[0] for v in agg.values()))
    import json
    with open(out.replace('_gemm_dram.md', '').rsplit('/', 1)[0] + '/gemm_traffic.json', 'w') as f:
        json.dump({'dram_bytes_per_launch': per_launch, 'launches': sum(v[0] for v in agg.values()), 'total_ms': T * 1e3,
                   'source': out, 'command': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... -k regex:gemm_tc python tools/ncu_step.py sr_tiny 256 2'}, f)
    return per_launch


if __name__ == '__main__':
    import os
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    if os.path.exists('gpurun_out/launches_%s.csv' % tag):
        agg, tot = launches('gpurun_out/launches_%s.csv' % tag, 'profiles/%s_launches.md' % tag)
        print('kernel time %.2f ms' % (tot / 1e3))
    if os.path.exists('gpurun_out/gemm_dram_%s.csv' % tag):
        per_launch = gemm_dram('gpurun_out/gemm_dram_%s.csv' % tag, 'profiles/%s_gemm_dram.md' % tag)
        print('gemm dram bytes per launch %.3e' % per_launch)
