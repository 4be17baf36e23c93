"""Headline metrics per kernel of an ncu report: python tests/tools/ncu_brief.py report.ncu-rep [kernel regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[0]; idx = {n: i for i, n in enumerate(h)}
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'launch__shared_mem_config_size']
stalls = [n for n in h if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio')]
for row in r[2:]:
    name = row[idx['Kernel Name']]
    if pat and not pat.search(name):
        continue
    print('----', name[:90])
    for w in want:
        if w in idx:
            print('  %-58s %s %s' % (w, row[idx[w]], r[1][idx[w]]))
    st = sorted(((float(row[idx[n]] or 0), n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for n in stalls), reverse=True)
    print('  stalls/issue:', ', '.join('%s %.2f' % (n, v) for v, n in st[:7]))
