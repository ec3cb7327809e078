"""Summarise an .ncu-rep (details page) into a small text file for profiles/."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'details', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
keep = ('Duration', 'DRAM Throughput', 'Memory Throughput', 'Compute (SM) Throughput', 'SM Frequency', 'DRAM Frequency',
        'Registers Per Thread', 'Theoretical Occupancy', 'Achieved Occupancy', 'Achieved Active Warps Per SM',
        'Issue Slots Busy', 'Executed Ipc Active', 'No Eligible', 'Eligible Warps Per Scheduler', 'Active Warps Per Scheduler',
        'L1/TEX Hit Rate', 'L2 Hit Rate', 'Mem Busy', 'Max Bandwidth', 'Mem Pipes Busy', 'Grid Size', 'Block Size',
        'Waves Per SM', 'Warp Cycles Per Issued Instruction', 'Avg. Active Threads Per Warp', 'Executed Instructions',
        'Dynamic Shared Memory Per Block', 'Static Shared Memory Per Block', 'Local Load', 'Elapsed Cycles', 'SM Active Cycles')
with open(out, 'w') as fh:
    name = None
    for r in rows[1:]:
        if len(r) < 15: continue
        if name != (r[0], r[4]):
            name = (r[0], r[4]); fh.write('\n== launch %s: %s  grid %s block %s\n' % (r[0], r[4][:90], r[7], r[8]))
        if r[12] in keep:
            fh.write('  %-28s %-44s %-14s %s\n' % (r[11][:28], r[12], r[13], r[14]))
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    for row in rr[2:]:
        fh.write('\n-- raw (launch %s)\n' % row[0])
        for i, h in enumerate(hdr):
            if h in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'smsp__inst_executed.sum',
                     'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
                     'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg') or (h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_warp_active.pct') ) or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
                fh.write('  %-95s %-10s %s\n' % (h, units[i], row[i]))
print(open(out).read())
