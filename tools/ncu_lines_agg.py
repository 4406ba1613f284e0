"""Per-CUDA-source-line summary aggregated over ALL launches of the kernels matching a regex in an .ncu-rep:
   python tools/ncu_lines_agg.py <report> <kernel regex> [top]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
def f(x):
    try: return float(x.replace(",", ""))
    except ValueError: return 0.0
hdr, fname = None, ""
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": hdr = None; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr and r[0].isdigit():
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [r[1], 0.0, 0.0])
        a[1] += f(r[hdr["# Samples"]]); a[2] += f(r[hdr["Instructions Executed"]])
ts, ti = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
print("samples", ts, "warp instructions", ti)
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{a[1] / ts * 100:5.1f}% smp {a[2] / ti * 100:5.1f}% inst  {fn}:{ln:<4d} | {a[0].strip()[:110]}")
