"""Turn `ncu -i <rep> --page raw --csv` into a compact per-launch table (markdown) + DRAM traffic totals."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, units, data = rows[0], rows[1], rows[2:]
c = lambda n: h.index(n)
EPI = {'0': 'plain', '1': 'semch', '2': 'global'}
cols = [('gpu__time_duration.sum', 'us', 1e-3), ('sm__cycles_elapsed.max', 'cycles', 1),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %', 1),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %', 1),
        ('dram__bytes_read.sum', 'rd MB', 1), ('dram__bytes_write.sum', 'wr MB', 1),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %', 1), ('launch__registers_per_thread', 'regs', 1),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %', 1)]
cols = [x for x in cols if x[0] in h]


def val(r, name):
    v = float(r[c(name)].replace(',', ''))
    u = units[c(name)]
    if u == 'Gbyte':
        v *= 1000
    elif u == 'Kbyte':
        v /= 1000
    elif u == 'byte':
        v /= 1e6
    elif u == 'ns':
        v /= 1000
    elif u == 'ms':
        v *= 1000
    return v


print('| # | kernel | ' + ' | '.join(x[1] for x in cols) + ' | DRAM GB/s |')
print('|---|---|' + '---|' * (len(cols) + 1))
tot_r = tot_w = tot_t = 0.0
fam_r = fam_w = 0.0
fam_n = 0
for i, r in enumerate(data):
    kn = r[c('Kernel Name')]
    if 'gemm_tc_kernel' in kn:
        epi = 'gemm ' + EPI.get(kn[kn.index('<') + 1:].replace('(int)', '').strip()[0], '?')
    else:
        epi = next((n for n in ('expand', 'rowdot', 'global_mix', 'shrink') if n in kn), kn[:16])
    fam = epi.startswith('gemm') or epi == 'global_mix'
    if fam:
        fam_r += val(r, 'dram__bytes_read.sum'); fam_w += val(r, 'dram__bytes_write.sum'); fam_n += 1
    vals = [val(r, x[0]) for x in cols]
    tot_t += val(r, 'gpu__time_duration.sum')
    tot_r += val(r, 'dram__bytes_read.sum')
    tot_w += val(r, 'dram__bytes_write.sum')
    gbs = (val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')) / val(r, 'gpu__time_duration.sum') * 1e3   # MB/us -> GB/s
    print('| %d | %s | ' % (i, epi) + ' | '.join(('%.0f' % v) if v >= 100 else ('%.1f' % v) for v in vals) + ' | %.0f |' % gbs)
print()
print(json.dumps({'launches': len(data), 'sum_duration_us': tot_t, 'dram_read_MB': tot_r, 'dram_write_MB': tot_w,
                  'dram_bytes_per_step_all_kernels': (tot_r + tot_w) * 1e6,
                  'gemm_family_launches': fam_n, 'dram_bytes_per_step': (fam_r + fam_w) * 1e6,
                  'per_launch_mean': (fam_r + fam_w) * 1e6 / max(fam_n, 1)}))
