"""Turns an .ncu-rep (captured on the B200 box with `ncu --set full --clock-control none
--import-source on`, see /opt/skills/guides/B200_PROFILING.md) into the small text summary that is
committed next to it.  Usage: python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/out.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
STALLS = ["long_scoreboard", "no_instruction", "wait", "short_scoreboard", "math_pipe_throttle", "branch_resolving",
          "lg_throttle", "mio_throttle", "dispatch_stall", "barrier", "not_selected", "membar", "drain", "imc_miss", "sleeping"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(out, "w") as f:
        f.write("# summary of %s (ncu --set full --clock-control none; cold-cache, serialised replays)\n" % rep.split("/")[-1])
        for r in rows[2:]:
            f.write("\nkernel: %s\n" % r[hdr.index("Kernel Name")])
            for k in KEYS:
                if k in hdr:
                    f.write("  %-70s %s %s\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
            f.write("  warp stall reasons (avg warps stalled per issue-active cycle):\n")
            for s_ in STALLS:
                k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s_
                if k in hdr:
                    f.write("    %-22s %s\n" % (s_, r[hdr.index(k)]))
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if k in hdr:
                    tot += float(r[hdr.index(k)]) * mult.get(units[hdr.index(k)], 1.0)
            f.write("  dram traffic per launch (read+write): %.0f bytes\n" % tot)
            traffic[r[hdr.index("Kernel Name")].split("(")[0]] = tot
    return traffic


if __name__ == "__main__":
    t = main(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3:   # also record dram bytes per launch for bench.py's roofline.traffic
        import json, os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[sys.argv[3]] = {"dram_bytes_per_launch": max(t.values()), "kernel": max(t, key=t.get),
                          "source": os.path.basename(sys.argv[2])}
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
