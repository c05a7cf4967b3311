"""Turns the scratch ncu outputs of scripts/gpu_check.sh (gpurun_out/) into the tracked summaries under
profiles/:  <tag>_launches.csv (the launch list of the bench command), <tag>_step_kernel.md (selected
counters of the full capture of the step kernel) and traffic_per_launch.json (DRAM bytes per
replication-step, read by bench.py for roofline.traffic).
    usage: python scripts/summarize_profile.py <tag> [reps-in-the-full-capture]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
out, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(prof, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------------
src = os.path.join(out, f"launches_{tag}.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    with open(os.path.join(prof, f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 6 --warmup 3 --no-cpu-baseline --full-horizon-reps 0\n")
        f.write("# per-launch times are cold-cache and serialised: compare shares, not absolutes\n")
        f.write("id,kernel,grid,block,duration_ms\n")
        tot, by = 0.0, {}
        for r in rows:
            ms = float(r[-1]) / 1e6
            name = r[4].split("(")[0].replace("void ", "")
            f.write(f"{r[0]},{name},{r[8].strip('()').split(',')[0]},{r[7].strip('()').split(',')[0]},{ms:.4f}\n")
            tot += ms
            by[name] = by.get(name, 0.0) + ms
        f.write("# share of device time by kernel: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(by.items(), key=lambda kv: -kv[1])) + "\n")
    print("wrote", f"profiles/{tag}_launches.csv")

# ---- full captures of the three kernels of a step ---------------------------------------------------------
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__icc_request_hit_rate.pct",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
lines = [f"# {tag}: ncu --set full --clock-control none, one launch of each kernel of a step", "",
         f"Command: `FSB_BATCHES=1 python scripts/prof_step.py --reps {reps} --steps 3 --t0 300` (default Params shape, "
         "probabilistic selector).  Numbers under a profiler are not bench values.", ""]
traffic = {}
for kname in ("draws", "events1", "pass", "choice", "events2"):
    rep = os.path.join(out, f"prof_{tag}_{kname}.ncu-rep")
    if not os.path.exists(rep):  # the fixed-shape build of the event kernels (fsb_step_fixed.cu)
        rep = os.path.join(out, f"prof_{tag}_fix_{kname}.ncu-rep")
        if not os.path.exists(rep):
            continue
        kname_shown = "fix_" + kname
    else:
        kname_shown = kname
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    lines += [f"## fsb_{kname_shown}_kernel", "", "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in hdr:
            lines.append(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
            try:
                st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    lines.append("| warp stalls per issue (top) | " + ", ".join(f"{h} {v:.2f}" for v, h in sorted(st, reverse=True)[:6]) + " | |")

    def num(key):
        i = hdr.index(key)
        v = float(r[i].replace(",", ""))
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}.get(units[i].lower(), 1.0)
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    traffic[kname] = dram / reps
    inst = float(r[hdr.index("smsp__inst_executed.sum")].replace(",", ""))
    lines += ["", f"DRAM bytes per replication-step: {dram / reps / 1e3:.1f} KB; warp instructions per replication-step: {inst / reps:.0f}", ""]
if traffic:
    open(os.path.join(prof, f"{tag}_step_kernels.md"), "w").write("\n".join(lines) + "\n")
    tpath = os.path.join(prof, "traffic_per_launch.json")
    keep = {}
    if os.path.exists(tpath):  # (the config-4 entry comes from its own capture: profiles/*_config4.md)
        keep = {k: v for k, v in json.load(open(tpath)).items() if k == "config4"}
    json.dump({"tag": tag, "kernel": "fsb_pass_kernel", "dram_bytes_per_replication_step": traffic.get("pass"),
               "all_kernels": traffic,
               "source": f"profiles/{tag}_step_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum over {reps} replication-steps per launch)",
               **keep},
              open(tpath, "w"), indent=1)
    print("wrote", f"profiles/{tag}_step_kernels.md", {k: round(v / 1e3, 1) for k, v in traffic.items()})

# ---- the bench line itself ---------------------------------------------------------------------
for name in (f"bench_{tag}.json", f"bench_ref_{tag}.json"):
    p = os.path.join(out, name)
    if os.path.exists(p) and os.path.getsize(p):
        open(os.path.join(prof, name), "w").write(open(p).read())
        print("copied", name)
