"""Summarise an .ncu-rep: per-kernel key metrics + top stall instructions.  Usage: python tools/ncu_summary.py rep [n_top]"""
import csv, subprocess, sys, re, collections, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_op_utc', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct']
idx = {h: i for i, h in enumerate(hdr)}
for k, r in enumerate(rows[2:]):
    print(f"==== [{k}] {r[idx['Kernel Name']][:110]}")
    for w in want:
        for h in hdr:
            if h == w:
                print(f"   {h:75s} {units[idx[h]]:14s} {r[idx[h]]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) < 3: continue
    sh = srows[1]
    try:
        i_src = sh.index('Source'); i_s = sh.index('Warp Stall Sampling (All Samples)'); i_ex = sh.index('Instructions Executed')
    except ValueError:
        continue
    seen = set(); data = []
    for q in srows[2:]:
        if len(q) > i_ex and q[i_s].isdigit() and q[0] not in seen:
            seen.add(q[0]); data.append((int(q[i_s]), q[i_src].strip(), int(q[i_ex])))
    tot = sum(d[0] for d in data) or 1
    print(f"   total samples {tot}, warp-instr executed {sum(d[2] for d in data)}")
    for s_, src_, ex in sorted(data, reverse=True)[:ntop]:
        print(f"      {100*s_/tot:5.1f}%  ex={ex:8d}  {src_[:100]}")
