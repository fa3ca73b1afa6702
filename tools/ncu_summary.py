"""Developer tool: the metrics DESIGN.md quotes, out of an ncu capture (a .ncu-rep, or the CSV of `ncu -i X --page raw --csv`).

    python tools/ncu_summary.py <label> <capture> [<label> <capture> ...] > profiles/<name>.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__icc_request_hit_rate.pct', 'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum']
out = {}
args = sys.argv[1:]
for label, path in zip(args[0::2], args[1::2]):
    raw = open(path).read() if path.endswith('.csv') else subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, units, vals = rows[h], rows[h + 1], rows[h + 2]
    d = {'kernel': vals[hdr.index('Kernel Name')][:60]}
    for k in KEYS:
        if k in hdr:
            d[k] = [vals[hdr.index(k)], units[hdr.index(k)]]
    for i, k in enumerate(hdr):
        if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and 'not_issued' not in k:
            v = float(vals[i].replace(',', '') or 0)
            if v >= 0.05:
                d.setdefault('stalls_per_issue_active', {})[k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = round(v, 3)
    out[label] = d
print(json.dumps(out, indent=1))
