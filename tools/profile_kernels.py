#!/usr/bin/env python
"""tools/profile_kernels.py <capture.ncu-rep> <out_dir> <prefix> "<workload note>" [units=<n> unit=<name>]

Reads one `ncu --set full --import-source on` capture (any number of launches) through `ncu -i ... --page raw --csv` and
`--page source --csv`, and writes for every launch of a library kernel `<out_dir>/<prefix>_<kernel>[_k].json`:
duration, DRAM bytes (read / written), issue-slot use, warps per sub-partition, registers, pipe utilisation, the stall
reasons per issued instruction (top ones first) and the dynamic SASS opcode histogram (warp instructions by opcode).
These are the tracked summaries DESIGN.md and bench.py's roofline objects quote."""
import csv
import io
import json
import os
import re
import subprocess
import sys

STALLS = ["wait", "not_selected", "math_pipe_throttle", "long_scoreboard", "short_scoreboard", "barrier", "dispatch_stall", "no_instruction",
          "branch_resolving", "mio_throttle", "lg_throttle", "membar", "tex_throttle", "sleeping", "drain", "imc_miss", "misc", "selected"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "msecond": 1.0, "us": 1e-3, "usecond": 1e-3, "ns": 1e-6,
         "nsecond": 1e-6, "s": 1e3, "second": 1e3}


def short_name(full):
    m = re.search(r"(gc_\w+|adx_\w+|hca_\w+|wave_\w+|dsp_\w+|\w*interleave\w*)(<[^>]*>)?", full)
    if not m:
        return None
    name = m.group(1)
    if m.group(2):
        arg = re.sub(r"[^0-9a-zA-Z]+", "", m.group(2).replace("(int)", "").replace("(bool)", ""))
        name += "_" + arg
    return name


def opcode_histogram(rep, index):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(index), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_at = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
    if hdr_at is None:
        return None
    hdr = rows[hdr_at]
    i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
    ops, total = {}, 0
    for r in rows[hdr_at + 1:]:
        try:
            n = int(r[i_exec])
        except (ValueError, IndexError):
            continue
        tok = r[i_src].split()
        if tok and tok[0].startswith("@"):
            tok = tok[1:]
        if not tok:
            continue
        op = tok[0].rstrip(";").split(".")[0]
        ops[op] = ops.get(op, 0) + n
        total += n
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:16]
    return {"warp_instructions_in_source_page": total, "share_pct": {k: round(100.0 * v / total, 2) for k, v in top}} if total else None


def main():
    rep, out_dir, prefix, note = sys.argv[1:5]
    extra = dict(a.split("=", 1) for a in sys.argv[5:])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    os.makedirs(out_dir, exist_ok=True)
    seen = {}
    for index, r in enumerate(rows[2:]):
        d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
        name = short_name(d["Kernel Name"][1])
        if not name:
            continue

        def num(key, default=None):
            if key not in d or d[key][1] in ("", "n/a"):
                return default
            u, v = d[key]
            try:
                return float(v.replace(",", "")) * SCALE.get(u, 1.0)
            except ValueError:
                return default

        stalls = {}
        for s in STALLS:
            v = num(f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio")
            if v:
                stalls[s] = round(v, 3)
        rd, wr = num("dram__bytes_read.sum", 0.0), num("dram__bytes_write.sum", 0.0)
        ms = num("gpu__time_duration.sum")
        inst = num("smsp__inst_executed.sum", 0.0)
        out = {"kernel": name, "full_name": d["Kernel Name"][1], "workload": note,
               "source": f"ncu --set full --clock-control none --import-source on, launch {index} of {os.path.basename(rep)} (cold caches, serialised: shares, not absolutes)",
               "gpu__time_duration_ms": ms, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr,
               "dram_GB_per_s": round((rd + wr) / ms / 1e6, 1) if ms else None,
               "dram_throughput_pct_of_peak": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               "warp_instructions": inst, "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warps_per_smsp": num("smsp__warps_active.avg.per_cycle_active"),
               "pipe_alu_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
               "pipe_fma_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
               "pipe_fp64_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               "pipe_lsu_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
               "registers_per_thread": int(num("launch__registers_per_thread", 0)), "grid": int(num("launch__grid_size", 0)),
               "block": int(num("launch__block_size", 0)), "shared_mem_per_block": num("launch__shared_mem_per_block_static", 0.0) + num("launch__shared_mem_per_block_dynamic", 0.0),
               "stalls_per_issued_instruction": dict(sorted(stalls.items(), key=lambda kv: -kv[1])),
               "top_stall": max((k for k in stalls if k != "selected"), key=lambda k: stalls[k], default=None)}
        if "units" in extra and float(extra["units"]) > 0:
            out[f"warp_instructions_per_{extra.get('unit', 'unit')}"] = round(inst / float(extra["units"]), 2)
        out["sass_opcodes"] = opcode_histogram(rep, index)
        k = seen.get(name, 0)
        seen[name] = k + 1
        path = os.path.join(out_dir, f"{prefix}_{name}{'' if k == 0 else '_' + str(k)}.json")
        json.dump(out, open(path, "w"), indent=1)
        print(path, round(ms, 3), "ms", "issue", out["issue_active_pct"], "top stall", out["top_stall"])


if __name__ == "__main__":
    main()
