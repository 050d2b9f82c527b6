"""Turn gpurun_out/ ncu captures into the committed summaries under profiles/.
    python scripts/summarize_profiles.py launches <launches.csv> <out.csv> "<header comment>"
    python scripts/summarize_profiles.py full <report.ncu-rep> <out.txt> "<header comment>"
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst, note):
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e6 if u in ('nsecond', 'ns') else v / 1e3 if u in ('usecond', 'us') else v
        a = agg.setdefault(row['Kernel Name'], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# {note}\nkernel,launches,total_ms,share\n')
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f'"{k[:110]}",{a[0]},{a[1]:.3f},{a[1] / tot:.4f}\n')
        f.write(f'TOTAL,,{tot:.3f},1\n')


def full(rep, dst, note):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = {h: (v, u) for h, v, u in zip(rows[0], rows[2], rows[1])}
    want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
            'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
            'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum']
    out = [f'# {note}']
    out += [f'{k:78s} {d[k][0]:>18s} {d[k][1]}' for k in want if k in d]
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, []
    for r in csv.reader(src.splitlines()):
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif len(r) >= 2 and r[0] == 'Line No':
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].strip().isdigit():
            try:
                agg.append((int(r[hdr.index('# Samples')]), cur, int(r[0]), r[1].strip()[:90], r))
            except ValueError:
                pass
    tot = sum(a[0] for a in agg) or 1
    names = ['stall_long_sb', 'stall_barrier', 'stall_short_sb', 'stall_wait', 'stall_selected', 'stall_not_selected', 'stall_math', 'stall_mio',
             'stall_lg', 'stall_branch_resolving', 'stall_no_inst', 'stall_dispatch']
    out += ['', '# warp-stall samples by reason (share of all samples)']
    for n in names:
        i = hdr.index(n)
        out.append(f'{n:26s} {sum(int(a[4][i] or 0) for a in agg) / tot * 100:5.1f}%')
    out += ['', '# samples by source file']
    by = collections.Counter()
    for a in agg:
        by[a[1]] += a[0]
    out += [f'{k:32s} {v / tot * 100:5.1f}%' for k, v in by.most_common()]
    out += ['', '# hottest source lines (share of samples, dominant stall)']
    for s_, f_, l_, txt, r in sorted(agg, key=lambda a: -a[0])[:16]:
        top = max(names, key=lambda n: int(r[hdr.index(n)] or 0))
        out.append(f'{s_ / tot * 100:5.1f}%  {f_}:{l_:<4d} {top:16s} | {txt}')
    open(dst, 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:5])
