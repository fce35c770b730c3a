"""Per-CUDA-source-line stall samples of one launch of an .ncu-rep.  Usage: python tools/ncu_lines.py rep launch_index [n_top]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2]); ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; fname = ""
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; i_s = r.index("Warp Stall Sampling (All Samples)"); i_ex = r.index("Instructions Executed"); continue
    if hdr is None or len(r) <= i_ex or not r[0]: continue   # SASS rows have an empty line number
    try: data.append((int(r[i_s]), fname, r[0], r[1].strip()[:120], r[i_ex]))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1
print("total samples", tot)
for s, f, ln, src, ex in sorted(data, reverse=True)[:ntop]:
    print(f"{100*s/tot:5.1f}%  {f}:{ln:>4s} ex={ex:>9s}  {src}")
