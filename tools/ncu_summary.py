"""Summaries of ncu exports for profiles/:
   python tools/ncu_summary.py launches <launches.csv>            per-launch timeline of the LAST call + totals by kernel
   python tools/ncu_summary.py full <report.ncu-rep> <out.csv>    per-kernel table of a --set full capture"""
import collections
import csv
import subprocess
import sys


def short(name):
    return name.split('(')[0].split('::')[-1]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, gs = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
    half = len(data) // 2
    acc = collections.OrderedDict()
    tot = 0.0
    for r in data[half:]:
        us = float(r[mv].replace(',', '')) / 1000.0
        tot += us
        a = acc.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += us
    for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print('%-22s %3d launches %9.1f us %5.1f %%' % (k, n, us, 100 * us / tot))
    print('total %.1f us, %d launches' % (tot, len(data) - half))


def full(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
            ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
            ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
            ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1_pct'),
            ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
            ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_pct'),
            ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
            ('sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'pipe_adu_pct'),
            ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'pipe_lsu_pct'),
            ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'pipe_alu_pct'),
            ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block')]
    cols = [(hdr.index(k), n) for k, n in want if k in hdr]
    with open(out, 'w') as f:
        f.write('kernel,' + ','.join('%s[%s]' % (n, units[i]) if units[i] else n for i, n in cols) + '\n')
        for r in rows[2:]:
            f.write(short(r[hdr.index('Kernel Name')]) + ',' + ','.join(r[i].replace(',', '') for i, _ in cols) + '\n')
    print(open(out).read())


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3])
