"""Turn the ncu outputs of scripts/gpu_profile.sh (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/summarise_profile.py r1
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_tex_mem_texture.sum", "l1tex__t_sectors_pipe_tex_mem_texture.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(units, r))) for r in rows[2:]]


limiter = {}


def summarise_full(rep, name):
    recs = raw_page(rep)
    lines = [f"# ncu --set full --clock-control none: {os.path.basename(rep)} ({tag})", ""]
    traffic = {}
    for rec in recs:
        kname = rec["Kernel Name"][1]
        lines += [f"## {kname}", "", "| metric | unit | value |", "|---|---|---|"]
        for k in KEYS:
            if k in rec:
                lines.append(f"| {k} | {rec[k][0]} | {rec[k][1]} |")
        def gb(key):
            u, v = rec[key]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
            return float(v.replace(",", "")) * scale
        t = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        traffic[kname] = t
        pct = lambda key: round(float(rec[key][1].replace(",", "")), 1) if key in rec else None  # noqa: E731
        limiter[kname] = {
            "tex_writeback_active_pct": pct("l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "lanes_active_of_32": pct("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "l1_hit_pct": pct("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": pct("lts__t_sector_hit_rate.pct")}
        lines += ["", f"DRAM traffic (read + write) for this launch: {t / 1e9:.3f} GB", ""]
    open(os.path.join(out, f"{tag}_{name}_ncu_full.md"), "w").write("\n".join(lines))
    return traffic


def summarise_launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
        agg.setdefault(r[ki], []).append(t)
    tot = sum(sum(v) for v in agg.values())
    lines = [f"# ncu launch list (gpu__time_duration.sum, --clock-control none) of the timed region of bench.py ({tag})",
             "", f"total kernel time {tot:.3f} ms over {sum(len(v) for v in agg.values())} launches "
             "(cold-cache, serialised: compare shares, not absolutes)", "",
             "| share | total ms | launches | kernel |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| {100 * sum(v) / tot:.2f}% | {sum(v):.3f} | {len(v)} | `{k[:110]}` |")
    open(os.path.join(out, f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")


g = os.path.join(ROOT, "gpurun_out")
if os.path.exists(os.path.join(g, f"{tag}_launches.csv")):
    summarise_launches(os.path.join(g, f"{tag}_launches.csv"))
tr = {}
# (report, summary name, C-ABI entry point whose launch it is, poses in the captured launch)
for rep, name, entry, batch in ((f"{tag}_prof_trilinear_fwd.ncu-rep", "trilinear_fwd", "xvr_trilinear_drr_fwd", 116),
                                (f"{tag}_prof_siddon_fwd.ncu-rep", "siddon_fwd", "xvr_siddon_drr_fwd", 32),
                                (f"{tag}_prof_staged.ncu-rep", "trilinear_staged", "xvr_trilinear_drr_fwd_staged", 116)):
    if os.path.exists(os.path.join(g, rep)):
        t = summarise_full(os.path.join(g, rep), name)
        top = max(t, key=t.get)
        tr[entry] = {"dram_bytes_per_launch": t[top], "batch": batch, "limiter": limiter.get(top),
                     "source": f"profiles/{tag}_{name}_ncu_full.md (ncu --set full capture, not the timed run)"}
for rep, name in ((f"{tag}_prof_volgrad.ncu-rep", "volume_grad_brick"), (f"{tag}_prof_siddon_volgrad.ncu-rep", "siddon_volume_grad")):
    if os.path.exists(os.path.join(g, rep)):
        summarise_full(os.path.join(g, rep), name)
if os.path.exists(os.path.join(g, "kernels.log")):
    body = open(os.path.join(g, "kernels.log")).read()
    table = body[body.index("| kernel |"):] if "| kernel |" in body else body
    open(os.path.join(out, f"{tag}_kernels.md"), "w").write(
        f"# scripts/bench_kernels.py on 1xB200 ({tag}): CUDA-event time per call, algorithmic bytes per SURVEY 8(d)\n\n"
        "Configs: trilinear = C2 (512^3, 256^2, B=116, n=500); siddon = C5 geometry (768^3, 512^2) at B=32; "
        "NCC = 116 images of 256^2.\n\n" + table)
if tr:
    path = os.path.join(out, "traffic.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(tr)
    json.dump(old, open(path, "w"), indent=1)
for f in ("bench_N1.json", "bench_ref.json"):
    if os.path.exists(os.path.join(g, f"{tag}_{f}")):
        open(os.path.join(out, f"{tag}_{f}"), "w").write(open(os.path.join(g, f"{tag}_{f}")).read())
print(os.listdir(out))
