"""profiles/ncu_traffic.json from an `ncu --set full` capture: DRAM bytes per launch of the dominant kernels (the LAST launch
of each kernel in the capture; entries of earlier captures for other kernels are kept)"""
import csv, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[0]
kn, rd, wr, tm = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("gpu__time_duration.sum")
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in rows[2:]:
    name = r[kn]
    key = "pcg_pipe_kernel<0>" if "pcg_pipe_kernel<1, 0" in name or "pcg_pipe_kernel<2, 0" in name else \
          "pcg_pipe_kernel<1>" if "pcg_pipe_kernel" in name else \
          "pcg_fused_kernel<FtCfgD>" if "pcg_fused_kernel" in name and "1, 5>" in name else \
          "pcg_fused_kernel" if "pcg_fused_kernel" in name else name.split("(")[0].replace("void ", "")
    b = float(r[rd]) * unit[rows[1][rd]] + float(r[wr]) * unit[rows[1][wr]]
    res[key] = {"dram_bytes_per_launch": b, "read": float(r[rd]) * unit[rows[1][rd]], "write": float(r[wr]) * unit[rows[1][wr]],
                "duration_us_under_ncu": float(r[tm]), "kernel": name[:120], "source": sys.argv[1]}
try:
    old = json.load(open(sys.argv[2]))          # entries of other captures are kept
except Exception:
    old = {}
old.update(res)
res = old
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1))
