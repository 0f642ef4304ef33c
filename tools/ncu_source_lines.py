"""Aggregate an ncu report's source page by CUDA source line: python tools/ncu_source_lines.py <rep> [top]
Prints, per (file, line): warp instructions executed, share, stall samples."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.OrderedDict(); cur_file = None; hdr = None; first_kernel = None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        if first_kernel is None: first_kernel = row[1]
        elif row[1] != first_kernel and not agg_done: pass
        continue
    if row[0] == "Line No": hdr = row; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or row[0] == "" or not row[0].isdigit(): continue
    key = (cur_file, int(row[0]))
    try: n = int(row[ie]); s = int(row[isamp])
    except ValueError: continue
    a = agg.setdefault(key, [0, 0, row[1]]); a[0] += n; a[1] += s
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print(f"total warp-inst {tot}  samples {tots}  (all captured launches of the report)")
for (f, l), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot:5.2f}% inst {100*s/max(tots,1):5.2f}% smp  {f}:{l}  {src.strip()[:110]}")
