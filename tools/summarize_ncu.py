"""Turns the ncu outputs of tools/gpu_capture.sh (gpurun_out/<tag>_<cfg>_*) into the tracked summaries under profiles/:
  profiles/<tag>_<cfg>_launches.csv          the per-launch gpu__time_duration list (copied) + kernel shares printed
  profiles/<tag>_<cfg>_ncu_full_summary.csv  selected metrics of the --set full capture, one column per launch
  profiles/traffic.json                      dram bytes (read + write) per launch and kernel, read by bench.py
usage: python tools/summarize_ncu.py <tag> <cfg>
"""
import csv
import json
import re
import os
import shutil
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_selected_per_warp_active.pct"]


_TC_MODES = {"0": "comm", "1": "dec", "2": "enc", "3": "head"}


def _kname(full):
    """cm::policy_tc_kernel<4, 0>(cm::TcArgs) -> policy_tc_kernel<comm>; other kernels lose their template arguments.
    The mode argument of policy_tc_kernel is kept: the encoder and head launches of the large-team pipeline are different
    kernels with different traffic and must not collide in traffic.json."""
    head = full.split("(")[0]
    base = re.sub(r"<.*>", "", head).split("::")[-1].replace("void ", "").strip()
    if base == "policy_tc_kernel":
        m = re.search(r"<\s*\d+\s*,\s*(\d+)\s*>", head)
        if m:
            return f"{base}<{_TC_MODES.get(m.group(1), m.group(1))}>"
    return base


def main(tag, cfg):
    out = os.path.join(ROOT, "gpurun_out")
    prof = os.path.join(ROOT, "profiles")
    # ---- launch list ----
    src = os.path.join(out, f"{tag}_{cfg}_launches.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(prof, f"{tag}_{cfg}_launches.csv"))
        tot = defaultdict(float); cnt = defaultdict(int)
        with open(src) as f:
            rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        for r in rows[1:]:
            k = _kname(r[ki])
            tot[k] += float(r[vi].replace(",", "")); cnt[k] += 1
        s = sum(tot.values())
        for k in tot:
            print(f"{k:28s} launches {cnt[k]:4d}  mean {tot[k] / cnt[k] / 1e3:9.3f} us  share {tot[k] / s:6.1%}")
    # ---- full capture ----
    src = os.path.join(out, f"{tag}_{cfg}_raw.csv")
    if os.path.exists(src):
        with open(src) as f:
            rows = list(csv.reader(f))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        lines = [["metric", "unit"] + [f"launch{i}" for i in range(len(data))]]
        for key in ["Kernel Name", "Block Size", "Grid Size"] + KEEP:
            if key in col:
                lines.append([key, units[col[key]]] + [d[col[key]] for d in data])
        with open(os.path.join(prof, f"{tag}_{cfg}_ncu_full_summary.csv"), "w", newline="") as f:
            csv.writer(f).writerows(lines)
        tr = {}
        for d in data:
            k = _kname(d[col["Kernel Name"]])
            def val(name):
                v = float(d[col[name]].replace(",", "")); u = units[col[name]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            tr[k] = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
        tpath = os.path.join(prof, "traffic.json")
        allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
        allt["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture made by "
                         "tools/gpu_capture.sh (profiles/<tag>_<cfg>_ncu_full_summary.csv); read by bench.py for roofline.traffic")
        allt[cfg] = tr
        allt.setdefault("_source", {})[cfg] = f"{tag}_{cfg}_ncu_full_summary.csv"
        json.dump(allt, open(tpath, "w"), indent=1)
        print("traffic", tr)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "c2")
