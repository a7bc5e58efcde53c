"""Turns the raw ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.
usage: python scripts/ncu_summarize.py <round-tag> <launches.csv> <full.ncu-rep> <kernel regex> <steps in the capture>"""
import collections, csv, re, subprocess, sys
tag, launches, rep, kre, nsteps = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    agg[name][0] += 1; agg[name][1] += v; tot += v
with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --n1d 256 ({nsteps} passes of the path in the capture)\n")
    f.write(f"# cold-cache, serialised launch times: compare SHARES, not absolutes.  total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
    f.write("%-44s %8s %12s %10s %8s\n" % ("kernel", "launches", "total ms", "ms/pass", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-44s %8d %12.3f %10.3f %7.1f%%\n" % (k[:44], n, t, t / nsteps, 100 * t / tot))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
with open(f"profiles/{tag}_{kre}_ncu_full.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kre}, python bench.py --n1d 256; values per launch\n")
    for r in rows[2:]:
        f.write(f"## launch id {r[0]}: {r[hdr.index('Kernel Name')][:60]}\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w); f.write("  %-92s %18s %s\n" % (w, r[i], units[i]))
print("written")
