"""Print one line per profiled launch from an `ncu --csv` log: python scripts/ncu_table.py file.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
iK, iM, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows:
    try: d.setdefault((r[iID], r[iK]), {})[r[iM]] = float(r[iV].replace(',', ''))
    except ValueError: pass
short = {'gpu__time_duration.sum': 't', 'dram__bytes_read.sum': 'dramR', 'dram__bytes_write.sum': 'dramW', 'lts__t_sector_hit_rate.pct': 'L2hit%', 'l1tex__t_sector_hit_rate.pct': 'L1hit%',
         'lts__t_sectors_op_read.sum': 'ltsRdSect', 'lts__t_sectors_srcunit_tex_op_read.sum': 'ltsRdTexSect', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum': 'l1Sect',
         'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum': 'l1HitSect', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum': 'l1Req', 'launch__grid_size': 'grid',
         'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'lts%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1%',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram%', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%', 'l1tex__data_pipe_lsu_wavefronts_mem_lg.sum': 'wavefronts'}
for (id_, k), m in d.items():
    name = k.split('::')[-1][:40]
    out = []
    for key, v in m.items():
        s = short.get(key, key)
        if 'Sect' in s: out.append('%s %.0fMB' % (s, v * 32 / 1e6))
        elif s in ('dramR', 'dramW'): out.append('%s %.0fMB' % (s, v / 1e6))
        elif s == 't': out.append('t %.0fus' % (v / 1e3 if v > 5e4 else v))
        elif s in ('l1Req', 'wavefronts'): out.append('%s %.1fM' % (s, v / 1e6))
        else: out.append('%s %.0f' % (s, v))
    print(id_, name, ' '.join(out))
