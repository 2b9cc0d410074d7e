"""Pull the judged metrics out of an `ncu --set full` report into a small JSON: python profiles/ncu_extract.py <report.ncu-rep> > profiles/<name>.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio']
STALL = 'smsp__average_warps_issue_stalled_'


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    res = {'kernel': vals[hdr.index('Kernel Name')], 'report': path.split('/')[-1], 'stall_cycles_per_issue': {}}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            res[h] = {'value': float(v.replace(',', '')) if v else None, 'unit': u}
        elif h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            res['stall_cycles_per_issue'][h[len(STALL):-len('_per_issue_active.ratio')]] = round(float(v), 3)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
