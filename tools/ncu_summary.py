"""Turn an .ncu-rep (ncu --set full) into a compact per-kernel table (markdown) for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof_final.ncu-rep > profiles/r01_ncu_final_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram rd'),
    ('dram__bytes_write.sum', 'dram wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'tensor ops %'),
    ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe %'),
    ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem(tc) %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('smsp__sass_inst_executed_op_utcmma.sum', 'UTCMMA inst'),
    ('sm__cycles_elapsed.avg', 'SM cycles'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__occupancy_limit_shared_mem', 'occ(smem)'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    name_i = h.index('Kernel Name')
    cols = [(h.index(k), label) for k, label in KEYS if k in h]
    print('| # | kernel | ' + ' | '.join(f'{label} [{units[i]}]' if units[i] else label for i, label in cols) + ' |')
    print('|---|---|' + '---|' * len(cols))
    for n, r in enumerate(rows[2:]):
        vals = []
        for i, _ in cols:
            try:
                v = float(r[i].replace(',', ''))
                vals.append(f'{v:.4g}')
            except ValueError:
                vals.append(r[i])
        print(f'| {n} | `{r[name_i][:70]}` | ' + ' | '.join(vals) + ' |')


if __name__ == '__main__':
    main(sys.argv[1])
