"""Summarise an `ncu --page raw --csv` export: one line per captured launch with the metrics DESIGN.md cites."""
import csv
import sys

WANT = [('gpu__time_duration.sum', 'dur'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor_inst'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('smsp__cycles_active.avg', 'cyc')]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n, units[hdr.index(k)]) for k, n in WANT if k in hdr]
    ki = hdr.index('Kernel Name')
    print('kernel | ' + ' | '.join('%s[%s]' % (n, u) for _, n, u in cols))
    for r in rows[2:]:
        print(r[ki].split('(')[0][-28:] + ' | ' + ' | '.join(r[i] for i, _, _ in cols))


if __name__ == '__main__':
    main(sys.argv[1])
