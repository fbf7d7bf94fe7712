"""Key roofline metrics of an .ncu-rep (one line per captured launch): python tools/ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'dur'),
    ('sm__cycles_elapsed.max', 'cycles'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('dram__bytes_read.sum', 'dram_rd'),
    ('dram__bytes_write.sum', 'dram_wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
    ('lts__t_sector_hit_rate.pct', 'l2hit%'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'xbar2sm'),
    ('l1tex__m_l1tex2xbar_write_bytes.sum', 'sm2xbar'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1%'),
    ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'tc_smem%'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
]


def summarize(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, zip(row, units)))
        name = d['Kernel Name'][0].split('(')[0][-48:]
        parts = []
        for k, label in KEYS:
            if k in d:
                v, u = d[k]
                try:
                    v = '%.4g' % float(v.replace(',', ''))
                except ValueError:
                    pass
                parts.append('%s=%s%s' % (label, v, (' ' + u) if u and u != '%' else ''))
        print(name, '|', ' '.join(parts))


if __name__ == '__main__':
    for p in sys.argv[1:]:
        summarize(p)
