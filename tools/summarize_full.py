"""Key per-launch metrics of an `ncu --set full` capture -> a small markdown table under profiles/ (the .ncu-rep itself stays in gpurun_out/)."""
import csv
import io
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'time'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'),
        ('smsp__inst_executed.sum', 'warp inst')]


def main(rep, out, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    iname = hdr.index('Kernel Name')
    with open(out, 'w') as f:
        f.write('# %s\n# source: ncu --set full --clock-control none --import-source on (%s), one row per captured launch\n\n' % (title, rep))
        f.write('| kernel | ' + ' | '.join('%s (%s)' % (n, units[i]) if units[i] else n for i, n in cols) + ' |\n')
        f.write('|---|' + '---:|' * len(cols) + '\n')
        for r in data:
            name = r[iname].split('(')[0].replace('void ', '').replace('vsx::<unnamed>::', '')[:48]
            f.write('| `%s` | ' % name + ' | '.join(r[i] for i, _ in cols) + ' |\n')
    print(open(out).read()[:3000])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3])
