#!/usr/bin/env python3
"""Turn ncu artefacts brought back in gpurun_out/ into small text summaries under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/prof_shade_r01.ncu-rep  profiles/r01_k_shade.txt
  python tools/ncu_summary.py list gpurun_out/launches_r01.csv        profiles/r01_launches.txt

`rep`  : one `ncu --set full` capture -> per-kernel key metrics (duration, DRAM bytes, throughput %, occupancy,
         registers, IPC, divergence) + warp-stall breakdown from the SASS page.
`list` : the `--metrics gpu__time_duration.sum` launch list -> per-kernel count / total / mean / share of the step.
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict, defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "derived__smsp__sass_thread_inst_executed_op_ffma_pred_on_x2", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_integer_pred_on.sum", "local_load_requests", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]


def ncu(*args):
    r = subprocess.run(["ncu", *args], capture_output=True, text=True)
    return r.stdout


def summarize_rep(path, out):
    raw = ncu("-i", path, "--page", "raw", "--csv")
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary of {path.split('/')[-1]} (per captured launch)", ""]
    for r in body:
        lines.append(f"== launch id {r[ix['ID']]}: {r[ix['Kernel Name']]}  grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}")
        for k in KEYS:
            if k in ix and r[ix[k]] != "":
                lines.append(f"   {k:<62} {r[ix[k]]} {units[ix[k]]}")
        try:
            rd = float(r[ix["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[ix["dram__bytes_write.sum"]].replace(",", ""))
            ur, uw = units[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_write.sum"]]
            lines.append(f"   traffic = dram read {rd} {ur} + write {wr} {uw}")
        except Exception:
            pass
        lines.append("")
    sass = ncu("-i", path, "--page", "source", "--csv", "--print-source", "sass")
    rows = list(csv.reader(io.StringIO(sass)))
    k = 0
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            ix = {h: j for j, h in enumerate(hdr)}
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            tot = OrderedDict((s, 0) for s in stalls)
            n_sass = samples = winst = tinst = 0
            i += 2
            while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
                r = rows[i]
                if len(r) == len(hdr):
                    n_sass += 1
                    samples += int(r[ix["# Samples"]])
                    winst += int(r[ix["Instructions Executed"]])
                    tinst += int(r[ix["Thread Instructions Executed"]])
                    for s in stalls:
                        tot[s] += int(r[ix[s]])
                i += 1
            lines.append(f"== SASS page, capture {k}: {name}")
            lines.append(f"   static SASS instructions {n_sass}; warp instructions executed {winst}; thread instructions {tinst}; "
                         f"avg active threads / warp instr {tinst / max(1, winst):.2f}")
            for s, v in sorted(tot.items(), key=lambda x: -x[1]):
                if v:
                    lines.append(f"   {s:<26} {100.0 * v / max(1, samples):5.1f} %")
            lines.append("")
            k += 1
        else:
            i += 1
    open(out, "w").write("\n".join(lines) + "\n")


def summarize_list(path, out):
    """Launch list (gpu__time_duration.sum, optionally dram__bytes_read.sum / dram__bytes_write.sum per launch)."""
    import json
    with open(path) as f:
        txt = [l for l in f if l.startswith('"')]
    rd = csv.reader(txt)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    scale_t = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_launch = {}
    order = []
    for r in rd:
        lid = r[ix["ID"]]
        name = r[ix["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        if lid not in per_launch:
            per_launch[lid] = {"name": name, "us": 0.0, "rd": 0.0, "wr": 0.0}
            order.append(lid)
        v = float(r[ix["Metric Value"]].replace(",", ""))
        m, u = r[ix["Metric Name"]], r[ix["Metric Unit"]]
        if m == "gpu__time_duration.sum":
            per_launch[lid]["us"] = v * scale_t.get(u, 1.0)
        elif m == "dram__bytes_read.sum":
            per_launch[lid]["rd"] = v * scale_b.get(u, 1.0)
        elif m == "dram__bytes_write.sum":
            per_launch[lid]["wr"] = v * scale_b.get(u, 1.0)
    agg = defaultdict(lambda: {"launches": 0, "us": 0.0, "max_us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    total = 0.0
    for lid in order:
        L = per_launch[lid]
        a = agg[L["name"]]
        a["launches"] += 1
        a["us"] += L["us"]
        a["max_us"] = max(a["max_us"], L["us"])
        a["dram_read"] += L["rd"]
        a["dram_write"] += L["wr"]
        total += L["us"]
    lines = [f"# launch list from {path.split('/')[-1]}: {len(order)} launches, {total / 1e3:.3f} ms of kernel time "
             f"(ncu --clock-control none; launches are serialised and cold-cache, so only the SHARES are comparable with bench.py)", "",
             f"{'kernel':<34}{'launches':>9}{'total ms':>12}{'mean us':>11}{'max us':>11}{'share':>8}{'DRAM rd MB':>13}{'DRAM wr MB':>13}{'DRAM B/launch':>15}"]
    for name, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
        lines.append(f"{name:<34}{a['launches']:>9}{a['us'] / 1e3:>12.3f}{a['us'] / a['launches']:>11.1f}{a['max_us']:>11.1f}"
                     f"{100 * a['us'] / total:>7.1f}%{a['dram_read'] / 1e6:>13.1f}{a['dram_write'] / 1e6:>13.1f}"
                     f"{(a['dram_read'] + a['dram_write']) / a['launches']:>15.0f}")
    open(out, "w").write("\n".join(lines) + "\n")
    json.dump({k: v for k, v in agg.items()}, open(out.rsplit(".", 1)[0] + ".json", "w"), indent=1)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    (summarize_rep if mode == "rep" else summarize_list)(src, dst)
    print(open(dst).read())
