"""Prints selected metrics of every launch in an .ncu-rep (reads `ncu -i rep --page raw --csv`)."""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        # the shared-memory data pipe, shared by the MMA operand fetch (tc) and the epilogue's loads / stores (lsu)
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'launch__waves_per_multiprocessor', 'dynamic_smem'] + sys.argv[2:]
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
for w in WANT:
    cols = [i for i, h in enumerate(hdr) if h == w or (w == 'dynamic_smem' and 'dynamic_shared' in h.lower())]
    if not cols:
        continue
    i = cols[0]
    vals = []
    for r in data:
        v = r[i]
        try:
            v = '%.4g' % float(v.replace(',', ''))
        except ValueError:
            v = v[:26]
        vals.append(v)
    print('| `%s` (%s) | %s |' % (hdr[i], units[i], ' | '.join(vals)))
