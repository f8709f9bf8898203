"""Summarise `ncu -i X.ncu-rep --page raw --csv` (one row per profiled launch) per kernel/grid -> markdown table:
duration, DRAM bytes read/written per launch, achieved DRAM GB/s, L2 hit rate, issue-slot and pipe utilisation, registers,
warps active.  usage: python profiles/ncu_raw_summary.py raw.csv [title]"""
import collections, csv, re, sys

COLS = [('us', 'gpu__time_duration.sum'), ('rd', 'dram__bytes_read.sum'), ('wr', 'dram__bytes_write.sum'),
        ('l2hit', 'lts__t_sector_hit_rate.pct'), ('issue', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        ('alu', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
        ('fma', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
        ('fp64', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
        ('tensor', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('warps', 'sm__warps_active.avg.pct_of_peak_sustained_active'), ('regs', 'launch__registers_per_thread')]
SCALE = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(path, title=''):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('<unnamed>::', '').replace('dmlmc::', '')
        key = (name, r[ix['Grid Size']], r[ix['Block Size']])
        for short, col in COLS:
            if col not in ix or r[ix[col]] in ('', 'n/a'):
                continue
            v = float(r[ix[col]].replace(',', '')) * SCALE.get(units[ix[col]], 1.0)
            agg[key][short].append(v)
    if title:
        print('## ' + title + '\n')
    print('| kernel | grid | block | launches | avg us | DRAM rd MB | DRAM wr MB | DRAM GB/s | L2 hit % | issue % | ALU % | FMA % | FP64 % | tensor % | warps % | regs |')
    print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
    for key, d in sorted(agg.items(), key=lambda kv: -sum(kv[1]['us'])):
        a = {k: sum(v) / len(v) for k, v in d.items()}
        g = lambda k, f='%.1f': (f % a[k]) if k in a else '-'
        gbs = (a.get('rd', 0) + a.get('wr', 0)) / a['us'] / 1e3
        print('| `%s` | %s | %s | %d | %.2f | %.2f | %.2f | %.0f | %s | %s | %s | %s | %s | %s | %s | %s |' % (
            key[0], key[1], key[2], len(d['us']), a['us'], a.get('rd', 0) / 1e6, a.get('wr', 0) / 1e6, gbs, g('l2hit'), g('issue'),
            g('alu'), g('fma'), g('fp64'), g('tensor'), g('warps'), g('regs', '%d')))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
