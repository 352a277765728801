"""Per-source-line view of one kernel of an .ncu-rep (needs -lineinfo and --import-source on):
warp-instructions executed, average active threads and stall samples per line of OUR sources.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top N] > profiles/NAME_lines.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # blocks: "File Path", "Function Name", header, then lines; the first kernel instance matching `want` is used
    per_line = defaultdict(lambda: [0, 0, 0, 0])   # inst, thread-inst, samples, long_sb samples
    seen_addr = set()
    cur_file, cur_fn, hdr, line_no, src = None, None, None, None, None
    fn_selected = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            cur_fn = r[1]
            if fn_selected is None and want in cur_fn:
                fn_selected = cur_fn
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_thr = hdr.index("Thread Instructions Executed")
            i_smp = hdr.index("# Samples")
            i_lsb = hdr.index("stall_long_sb")
            continue
        if hdr is None or cur_fn != fn_selected:
            continue
        if r[0] != "":
            line_no, src = r[0], r[1]
            continue
        addr = r[2]
        if addr in ("...", "-") or (cur_fn, addr) in seen_addr:
            continue
        seen_addr.add((cur_fn, addr))
        try:
            v = per_line[(cur_file, int(line_no), src)]
            v[0] += int(r[i_inst]); v[1] += int(r[i_thr]); v[2] += int(r[i_smp]); v[3] += int(r[i_lsb])
        except (ValueError, IndexError):
            pass
    tot_inst = sum(v[0] for v in per_line.values()) or 1
    tot_thr = sum(v[1] for v in per_line.values())
    tot_smp = sum(v[2] for v in per_line.values()) or 1
    print(f"# per-line profile of `{fn_selected}` ({rep})\n")
    print(f"warp instructions {tot_inst}, thread instructions {tot_thr}, mean active threads {tot_thr / tot_inst:.2f}, stall samples {tot_smp}\n")
    print("| file:line | inst % | active thr | samples % | long_sb % of line | source |")
    print("|---|---|---|---|---|---|")
    for (f, ln, s), v in sorted(per_line.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"| {f}:{ln} | {100 * v[0] / tot_inst:.1f} | {v[1] / max(1, v[0]):.1f} | {100 * v[2] / tot_smp:.1f} | "
              f"{100 * v[3] / max(1, v[2]):.0f} | `{s.strip()[:110]}` |")


if __name__ == "__main__":
    main()
