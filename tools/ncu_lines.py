"""aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H = [r for r in rows if 'Instructions Executed' in r][0]
ie, smp, te = H.index('Instructions Executed'), H.index('# Samples'), H.index('Thread Instructions Executed')
def num(s):
    try: return float(s)
    except ValueError: return None
data = [r for r in rows if len(r) > te and r[0] not in ('', 'Line No') and num(r[ie]) is not None]
tot = sum(num(r[ie]) for r in data); tots = sum(num(r[smp]) or 0 for r in data)
print("total warp inst %.4g  samples %d" % (tot, tots))
for r in sorted(data, key=lambda r: -(num(r[smp]) or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 22]:
    print("L%-4s inst %5.1f%%  samples %5.1f%%  thr/warp %5.1f | %s" % (r[0], 100 * num(r[ie]) / tot, 100 * (num(r[smp]) or 0) / tots,
          (num(r[te]) or 0) / max(num(r[ie]), 1), r[1].strip()[:100]))
