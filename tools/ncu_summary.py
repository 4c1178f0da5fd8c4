"""Summarise an .ncu-rep (first kernel): headline metrics, stall reasons, hottest SASS instructions."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
print("kernel:", m.get("Kernel Name", ("?",))[0][:100])
for k in keys:
    if k in m: print("  %-75s %s %s" % (k, m[k][0], m[k][1]))
st = [(float(v[0]), k.split("issue_stalled_")[1].split("_per_")[0]) for k, v in m.items()
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[0]]
print("  stalls per issue:", ", ".join("%s %.2f" % (n, x) for x, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
isrc, isamp, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
data = [(int(r[isamp]), r[isrc].strip(), int(r[iex]), i) for i, r in enumerate(rows[2:]) if len(r) > isamp and r[isamp].isdigit()]
tot = sum(d[0] for d in data)
g = collections.Counter(); ge = collections.Counter()
for s, t, ex, i in data:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    g[op] += s; ge[op] += ex
print("  samples by opcode:", ", ".join("%s %.1f%% (ex %d)" % (k, 100 * v / tot, ge[k]) for k, v in g.most_common(12)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for s, t, ex, i in sorted(data, reverse=True)[:n]:
    print("  %5.1f%% ex=%8d #%-5d %s" % (100 * s / tot, ex, i, t[:80]))
