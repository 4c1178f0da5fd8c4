"""profiles/propagate_ncu_facts.json from an ncu capture of the propagation kernel (config 2):
   python tools/ncu_facts.py gpurun_out/<capture>.ncu-rep "<how it was captured>"
DRAM bytes per launch (bench.py roofline.traffic) and the shared-memory pipe's utilisation, from which the on-chip
ceiling of this kernel design follows: at 100 % of the LDS pipe it would run 1 / lds_pct faster."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, how = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def num(k):
    v, u = m[k]
    x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(u, 1.0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
B, H, W, C, P = 16, 321, 321, 21, 48
alg = 4 * (P + 2 * C) * B * H * W
t_us = num("gpu__time_duration.sum")
rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
lds = num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
frac = alg / (t_us * 1e-6) / 1e9 / peak
out = {"kernel": m["Kernel Name"][0][:80], "config": "B=16, C=21, 321x321 (config 2), one launch of the tile kernel",
       "gpu_time_duration_us": t_us, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
       "algorithmic_bytes_per_launch": alg,
       "note": "ncu's write counter stops at kernel end: the part of the output still dirty in the 126 MB L2 is not in it",
       "hbm_frac_in_capture": frac, "lds_pct": lds, "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "lds_ceiling_frac_of_hbm_roofline": frac / (lds / 100.0),
       "source": "ncu --set full --clock-control none (" + how + "), " + os.path.basename(rep)}
json.dump(out, open(os.path.join(ROOT, "profiles", "propagate_ncu_facts.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
