"""Summarise ncu reports into profiles/: per-kernel key metrics (raw page) and the launch-list shares."""
import csv, subprocess, sys, collections
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
def raw(rep):
    # a .ncu-rep report, or the CSV that `ncu -i report --page raw --csv` printed (reports stay on the GPU box: size)
    out = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if not l.startswith('==')))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for k in KEYS:
            if k in hdr: d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        stalls = []
        for i, k in enumerate(hdr):
            if 'average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio'):
                try: stalls.append((float(r[i]), k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
                except ValueError: pass
        d['top_stalls(warps per issue)'] = ', '.join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:5])
        res.append(d)
    return res
def launches(path):
    agg = collections.OrderedDict()
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    for r in rows[1:]:
        if len(r) <= vi: continue
        v = float(r[vi].replace(',', '')); u = r[ui]
        ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v if u in ('ms', 'msecond') else v * 1e3
        a = agg.setdefault(r[ki][:90], [0, 0.0]); a[0] += 1; a[1] += ms
    return agg
if __name__ == '__main__':
    mode = sys.argv[1]
    if mode == 'raw':
        for d in raw(sys.argv[2]):
            print('## ' + d.pop('kernel'))
            for k, v in d.items(): print(f"- {k}: {v}")
            print()
    else:
        agg = launches(sys.argv[2]); tot = sum(a[1] for a in agg.values())
        print('| kernel | launches | total ms | share |\n|---|---|---|---|')
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| `{k}` | {c} | {ms:.3f} | {100*ms/tot:.1f}% |")
