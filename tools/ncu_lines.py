"""aggregate an `ncu --page source --csv --print-source cuda,sass` dump by (file, CUDA source line); optional line ranges
of one file summed as phases:  ncu_lines.py dump.csv [N] [file:lo-hi=name ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
fname, data, H = "", [], None
def num(s):
    try: return float(s)
    except ValueError: return None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r: H = r; continue
    if H is None or len(r) < len(H) or r[0] in ("", "Line No"): continue
    ie, smp, te = H.index("Instructions Executed"), H.index("# Samples"), H.index("Thread Instructions Executed")
    if num(r[ie]) is None: continue
    data.append((fname, int(r[0]), r[1].strip(), num(r[ie]), num(r[smp]) or 0, num(r[te]) or 0))
tot = sum(d[3] for d in data); tots = sum(d[4] for d in data)
print("total warp inst %.4g  samples %d" % (tot, tots))
N = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 22
for d in sorted(data, key=lambda d: -d[4])[:N]:
    print("%-22s L%-4d inst %5.1f%%  samples %5.1f%%  thr/warp %5.1f | %s" % (d[0][:22], d[1], 100 * d[3] / tot, 100 * d[4] / tots, d[5] / max(d[3], 1), d[2][:90]))
for spec in sys.argv[3:]:
    rng, name = spec.split("=")
    f, lh = rng.split(":"); lo, hi = [int(v) for v in lh.split("-")]
    sel = [d for d in data if d[0] == f and lo <= d[1] <= hi]
    print("phase %-28s inst %5.1f%%  samples %5.1f%%" % (name, 100 * sum(d[3] for d in sel) / tot, 100 * sum(d[4] for d in sel) / tots))
