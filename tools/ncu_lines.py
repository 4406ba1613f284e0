"""Per-CUDA-source-line summary (samples, executed instructions) of one kernel in an .ncu-rep:
   python tools/ncu_lines.py <report> <kernel regex> [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
def f(x):
    try: return float(x.replace(",", ""))
    except ValueError: return 0.0
fn, hdr, lines, fname = None, None, [], ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; hdr = None; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr and r[0] not in ("", "-") and r[0].isdigit():
        lines.append((fn, fname, int(r[0]), r[1], f(r[hdr["# Samples"]]), f(r[hdr["Instructions Executed"]])))
first = lines[0][0]
lines = [l for l in lines if l[0] == first]
print(first)
ts, ti = sum(l[4] for l in lines), sum(l[5] for l in lines)
print("samples", ts, "warp instructions", ti)
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{l[4] / ts * 100:5.1f}% smp {l[5] / ti * 100:5.1f}% inst  {l[1]}:{l[2]:<4d} | {l[3].strip()[:105]}")
