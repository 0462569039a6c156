"""Turns gpurun_out/*.ncu-rep / launch-list CSVs into the small text/JSON summaries committed under profiles/."""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def summarise(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                d[h] = f"{vals[i]} {units[i]}".strip()
            if "issue_stalled" in h and "ratio" in h and "not_issued" not in h:
                try:
                    if float(vals[i]) > 0.3:
                        d.setdefault("stalls_per_issue", {})[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(float(vals[i]), 2)
                except ValueError:
                    pass
        r = to_bytes(vals[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        w = to_bytes(vals[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        d["dram_traffic_bytes"] = r + w
        out.append(d)
    return out


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        k = r[4].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(a[1] for a in agg.values())
    return [{"kernel": k, "launches": n, "total_us": round(t / 1e3, 1), "avg_us": round(t / n / 1e3, 2), "share": round(t / tot, 4)}
            for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    data = summarise(src) if mode == "rep" else launches(src)
    json.dump(data, open(dst, "w"), indent=1)
    print(json.dumps(data, indent=1)[:3000])
