"""python tools/ncu_summarize.py <report.ncu-rep> <key> [summary.json]
Boil one `ncu --set full --import-source on` report down to the numbers DESIGN.md / bench.py quote: per-launch metric
averages over the captured launches of the step kernel, the warp-stall mix and the share of samples per source file."""
import collections, csv, io, json, os, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__cycles_active.avg", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
           "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = {"launches_captured": len(data), "kernel": data[0][hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            vals = [float(r[i].replace(",", "")) for r in data if r[i] not in ("", "n/a")]
            if vals:
                res[m] = {"value": sum(vals) / len(vals), "unit": units[i]}
    # source page: stall mix + share per file
    src = ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"])
    stall = collections.Counter(); files = collections.Counter(); cur = None; h = None; total = 0
    for row in csv.reader(io.StringIO(src)):
        if not row: continue
        if row[0] == "File Path": cur = os.path.basename(row[1]); continue
        if row[0] == "Line No":
            h = row; si = row.index("# Samples")
            cols = [(i, k) for i, k in enumerate(row) if k.startswith("stall_") and "Not Issued" not in k]
            continue
        if h is None or row[0] == "": continue
        try: int(row[0])
        except ValueError: continue
        try:
            n = int(row[si] or 0); add = [(k, int(row[i] or 0)) for i, k in cols]
        except (ValueError, IndexError):
            continue          # a source line whose text (inline asm with quotes / commas) broke the CSV columns
        files[cur] += n; total += n
        for k, v in add: stall[k] += v
    st = sum(stall.values()) or 1
    res["warp_stall_sampling"] = {"samples": total,
                                  "stall_share_pct": {k: round(100.0 * v / st, 1) for k, v in stall.most_common() if 100.0 * v / st >= 1.0},
                                  "samples_by_file_pct": {k: round(100.0 * v / max(total, 1), 1) for k, v in files.most_common() if v * 100 >= total}}
    print(json.dumps(res, indent=1))
    if out_path:
        d = json.load(open(out_path)) if os.path.exists(out_path) else {}
        d[key] = res
        json.dump(d, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
