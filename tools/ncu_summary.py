"""Turns an ncu report (--set full) / launch list into the small CSV + text tables kept under profiles/.

    python tools/ncu_summary.py full  gpurun_out/x.ncu-rep  profiles/out.csv     (needs ncu on PATH; no GPU)
    python tools/ncu_summary.py list  gpurun_out/launches.csv
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = (r'(Kernel Name|ID|launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit.*)|'
        r'gpu__time_duration.sum|dram__bytes_(read|write).sum|sm__throughput.avg.pct_of_peak_sustained_elapsed|'
        r'sm__inst_executed_pipe_(alu|fma|lsu|fmaheavy|fmalite|uniform|xu).avg.pct_of_peak_sustained_active|'
        r'smsp__issue_active.avg.per_cycle_active|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum|'
        r'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|'
        r'smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|'
        r'sm__cycles_active.avg|smsp__cycles_active.avg)$')


def short(name):
    m = re.search(r'(k_\w+)<([^>]*)>', name)
    if m:
        return m.group(1).replace('k_search8_cu', 'cu').replace('k_frac_tiles', 'frac').replace('k_search8', 's8') + '<' + \
            m.group(2).replace('unsigned char', 'u8').replace(' ', '') + '>'
    return re.sub(r'\(.*', '', name)[:24]


def full(rep, out_csv):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [i for i, h in enumerate(hdr) if re.match(KEEP, h)]
    with open(out_csv, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in keep])
        w.writerow([units[i] for i in keep])
        for r in rows[2:]:
            w.writerow([r[i] for i in keep])
    ki = hdr.index('Kernel Name')
    print('metric'.ljust(30) + ''.join(short(r[ki])[:13].rjust(14) for r in rows[2:]))
    for i in keep:
        if hdr[i] in ('Kernel Name', 'ID'):
            continue
        s = hdr[i].replace('smsp__average_warps_issue_stalled_', 'st_').replace('_per_issue_active.ratio', '') \
                  .replace('.avg.pct_of_peak_sustained_active', '%').replace('sm__inst_executed_pipe_', 'pipe_')
        vals = []
        for r in rows[2:]:
            try:
                vals.append(f"{float(r[i].replace(',', '')):14.3g}")
            except ValueError:
                vals.append(r[i][:13].rjust(14))
        print(s[:29].ljust(30) + ''.join(vals))


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1e3 if row['Metric Unit'] == 'ns' else v * 1e3 if row['Metric Unit'] == 'ms' else v
        a = agg.setdefault(short(row['Kernel Name']), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    for k, (n, t) in agg.items():
        print(f"{k:44s} n={n:4d} avg_us={t / n:10.1f} share={100 * t / tot:5.1f}%")


if __name__ == '__main__':
    if sys.argv[1] == 'full':
        full(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2])
