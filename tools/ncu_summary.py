"""Summarise `ncu --set full` captures (not a test, not shipped):

    python tools/ncu_summary.py <tag> <label>=<file.ncu-rep> [<label>=<file.ncu-rep> ...] [--traffic n256=<label> n4096=<label>]

writes profiles/<tag>_ncu_full_summary.json (selected launch-level metrics per capture) and, with --traffic, refreshes
profiles/traffic.json (DRAM bytes per launch next to the algorithmic bytes) stamped with the hash of the kernel sources
the capture was taken on -- bench.py only reports `roofline.traffic` when that hash matches the build it runs."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
UNIT_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_page(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    tag, args = sys.argv[1], sys.argv[2:]
    traffic = {}
    if "--traffic" in args:
        i = args.index("--traffic")
        traffic = dict(a.split("=") for a in args[i + 1:])
        args = args[:i]
    summary, dram = {}, {}
    for a in args:
        label, path = a.split("=")
        launches, units = raw_page(path)
        r = launches[-1]
        summary[label] = {m: {"value": r[m], "unit": units[m]} for m in METRICS if m in r}
        summary[label]["kernel"] = r.get("Kernel Name", "")
        dram[label] = sum(float(r[m].replace(",", "")) * UNIT_BYTES[units[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    with open(os.path.join(ROOT, "profiles", "%s_ncu_full_summary.json" % tag), "w") as f:
        json.dump(summary, f, indent=0)
    if traffic:
        import bench
        t = {"source_hash": bench.source_hash(), "source": "profiles/%s_ncu_full_summary.json (ncu --set full --clock-control none, one launch each)" % tag}
        for key, label in traffic.items():
            n = int(key[1:])
            t["dram_bytes_per_launch_" + key] = dram[label]
            t["algorithmic_bytes_" + key] = n * bench.BYTES_PER_SPEC
            t["kernel_" + key] = summary[label]["kernel"]
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump(t, f, indent=1)
    print(json.dumps({k: {"us": v["gpu__time_duration.sum"]["value"], "dram_MB": dram[k] / 1e6, "issue": v["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"]} for k, v in summary.items()}))


if __name__ == "__main__":
    main()
