"""Summarise an .ncu-rep: key metrics + the SASS instructions with the most stall samples."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct', 'launch__grid_size',
        'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'lts__t_bytes.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d['Kernel Name'][:70])
    for k in keys:
        if k in d: print('   ', k, d[k], u[h.index(k)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for n, hi in enumerate(his):
    hh = rows[hi]
    si = hh.index('Source'); ss = hh.index('Warp Stall Sampling (All Samples)')
    end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    data = [(int(r[ss]), idx, r[si]) for idx, r in enumerate(rows[hi + 1:end]) if len(r) > ss and r[ss].isdigit()]
    tot = sum(d[0] for d in data) or 1
    print('--- kernel', n, 'total samples', tot, 'instructions', len(data))
    for s_, idx, src_ in sorted(data, reverse=True)[:topn]:
        print(f'{s_:8d} {100*s_/tot:5.1f}%  #{idx:5d} {src_.strip()}')
