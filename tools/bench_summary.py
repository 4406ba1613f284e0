"""Print the key numbers of a bench.py log (last JSON line)."""
import json
import sys

for path in sys.argv[1:]:
    lines = [l for l in open(path) if l.startswith("{")]
    if not lines:
        print(path, "NO JSON LINE")
        print(open(path).read()[-3000:])
        continue
    d = json.loads(lines[-1])
    print(f"== {path}: n_gpus {d['n_gpus']} value {d['value']:.4g} {d['unit']} ms/step {d['ms_per_step']:.3f}", end="")
    if d.get("e2e"):
        print(f" e2e {d['e2e']['value']:.4g} h2d {d['e2e']['h2d_bytes_per_step']:.3g}", end="")
    print()
    for k in ("kernel_ms", "phase_ms"):
        if k in d:
            print("  ", k, {a: round(b, 3) for a, b in d[k].items()})
    for k in ("roofline", "roofline_far", "roofline_hbm", "roofline_direct"):
        if d.get(k):
            r = d[k]
            print(f"   {k}: {r['achieved']:.2f} {r['unit']} frac {r['frac']:.3f} kernel_ms {r.get('kernel_ms', 0):.3f}")
    if d.get("parity"):
        print("   parity", d["parity"]["ok"], d["parity"]["alpha_max_rel"], d["parity"]["F_max_rel"])
    if d.get("cpu_baseline"):
        print("   cpu", d["cpu_baseline"]["kind"], round(d["cpu_baseline"]["value"], 1), "cores", d["cpu_baseline"]["cores"])
    if "spectra_per_s" in d:
        print("   spectra/s", d["spectra_per_s"], d.get("spectrum_checks"))
    print("   clocks", d.get("clocks"))
