"""aggregate an `ncu --page source --csv --print-source sass` dump by barrier-delimited phase and list the hottest SASS lines"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if '# Samples' in r][0]
H = rows[hi]
si, ie, src, te = H.index('# Samples'), H.index('Instructions Executed'), H.index('Source'), H.index('Thread Instructions Executed')
wf, wfi = H.index('L1 Wavefronts Shared'), H.index('L1 Wavefronts Shared Ideal')
def f(x):
    try: return float(x)
    except ValueError: return 0.0
phase, acc, tot = 0, {}, 0
data = [r for r in rows[hi + 1:] if len(r) > te]
for r in data:
    a = acc.setdefault(phase, [0, 0, 0, 0, 0, 0])
    a[0] += f(r[si]); a[1] += f(r[ie]); a[2] += f(r[te]); a[3] += 1; a[4] += f(r[wf]); a[5] += f(r[wfi])
    tot += f(r[si])
    if 'BAR.SYNC' in r[src]: phase += 1
for p, a in acc.items():
    print("phase %d: samples %5.1f%%  warp-inst %.3g  thr/inst %.1f  sass lines %d  smem wavefronts %.3g (ideal %.3g)"
          % (p, 100 * a[0] / tot, a[1], a[2] / max(a[1], 1), a[3], a[4], a[5]))
data.sort(key=lambda r: -f(r[si]))
for r in data[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%6.2f%% inst %8s thr %5.1f wf %8s/%8s | %s" % (100 * f(r[si]) / tot, r[ie], f(r[te]) / max(f(r[ie]), 1), r[wf], r[wfi], r[src].strip()[:90]))
