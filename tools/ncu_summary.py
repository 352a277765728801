"""Summarises ncu --set full reports into profiles/: one markdown table per report and the machine-readable
profiles/traffic.json that bench.py reads (physical DRAM bytes per launch, issue statistics of the dominant kernels).

    python tools/ncu_summary.py <scene>=<report.ncu-rep> [...]      (run here, on the CPU box; needs only the ncu CLI)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
KEYS = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("smsp__inst_executed.sum", "inst_executed"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("dram__bytes_read.sum", "dram_read_bytes"), ("dram__bytes_write.sum", "dram_write_bytes"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch_resolving")]


def kind_of(name):
    if "k_tile<(int)1" in name or "k_tile<1" in name:
        return "k_tile"
    if "k_tile<(int)2" in name or "k_tile<2" in name:
        return "k_tile_ssaa"
    if "k_tile" in name:
        return "k_tile_queue"
    return name.split("(")[0].split("<")[0].replace("void ", "").replace("rtk::", "")


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"kernel": d["Kernel Name"][:70]}
        for k, short in KEYS:
            if k in d and d[k] != "":
                v = float(d[k].replace(",", ""))
                if short in ("us", "dram_read_bytes", "dram_write_bytes"):
                    v *= UNIT.get(u[k], 1.0)
                rec[short] = v
        rec["dram_bytes"] = rec.get("dram_read_bytes", 0.0) + rec.get("dram_write_bytes", 0.0)
        rec["kind"] = kind_of(d["Kernel Name"])
        launches.append(rec)
    return launches


def main():
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for arg in sys.argv[1:]:
        scene, rep = arg.split("=", 1)
        launches = read(rep)
        base = os.path.splitext(os.path.basename(rep))[0]
        md = [f"# {base}: ncu --set full --clock-control none ({scene}); per launch, caches cold per replay", "",
              "| metric | " + " | ".join(l["kind"] for l in launches) + " |", "|---|" + "---|" * len(launches)]
        for _, short in KEYS + [("", "dram_bytes")]:
            md.append(f"| {short} | " + " | ".join(f"{l.get(short, float('nan')):.6g}" for l in launches) + " |")
        sm, clk = 148, 1.965e9
        md.append("| issue floor us (inst / (148 SM x 4 x 1.965 GHz)) | " + " | ".join(f"{l['inst_executed'] / (sm * 4 * clk) * 1e6:.1f}" for l in launches) + " |")
        open(os.path.join(ROOT, "profiles", base + ".md"), "w").write("\n".join(md) + "\n")
        entry = traffic.setdefault(scene, {})
        for l in launches:
            e = entry.setdefault(l["kind"], {"launches": []})
            e["launches"] = [l]
            e["dram_bytes_per_launch"] = l["dram_bytes"]
            e["issue"] = {"inst_executed": l["inst_executed"], "issue_active_pct": l["issue_active_pct"], "threads_per_inst": l["threads_per_inst"],
                          "us_under_ncu": l["us"], "frac_of_issue_peak": l["inst_executed"] / (sm * 4 * clk) / (l["us"] * 1e-6)}
            e["source"] = f"profiles/{base}.md (ncu --set full, --clock-control none; caches cold per replay)"
        print(scene, [(l["kind"], round(l["us"], 1), int(l["dram_bytes"])) for l in launches])
    json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
