"""Summarise an .ncu-rep: key metrics + stall reasons per kernel, optionally the hottest SASS lines (--src N)."""
import csv, subprocess, sys
rep = sys.argv[1]
nsrc = int(sys.argv[sys.argv.index("--src") + 1]) if "--src" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'sm__cycles_elapsed.max ',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum ',
        'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')][:90])
    for w in want:
        for i, h in enumerate(hdr):
            if (h + ' ').startswith(w):
                print(f'  {h:75s} {r[i]} {rows[1][i]}')
    out = []
    for i, h in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try: out.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
    tot = sum(v for v, _ in out) or 1
    print('  stall samples: ' + ', '.join(f'{h} {100 * v / tot:.1f}%' for v, h in sorted(out, reverse=True)[:9]))
if nsrc:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = next(r for r in rows if 'Source' in r)
    ia, isamp, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    def I(x):
        try: return int(x)
        except ValueError: return 0
    data = [(I(r[isamp]), i, r[ia].strip(), I(r[iex])) for i, r in enumerate(rows) if len(r) > iex and r is not hdr]
    tot = sum(d[0] for d in data) or 1
    print(f'hottest SASS lines (of {tot} samples):')
    for s, i, t, ex in sorted(data, reverse=True)[:nsrc]:
        print(f'  {s:6d} {100 * s / tot:5.1f}%  line {i:5d} ex={ex:9d}  {t[:100]}')
