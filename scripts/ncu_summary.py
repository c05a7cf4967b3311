"""Key counters + stall breakdown of an ncu report (first kernel).  usage: ncu_summary.py <rep> [reps]"""
import csv, subprocess, sys
rep=sys.argv[1]; reps=int(sys.argv[2]) if len(sys.argv)>2 else 8192
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr,units,r=rows[0],rows[1],rows[2]
def g(k):
    return r[hdr.index(k)] if k in hdr else None
def num(k):
    v=float(g(k).replace(",","")); u=units[hdr.index(k)].lower()
    return v*{"gbyte":1e9,"mbyte":1e6,"kbyte":1e3,"byte":1.0,"tbyte":1e12,"msecond":1e-3,"usecond":1e-6,"second":1.0,"ms":1e-3,"us":1e-6}.get(u,1.0)
print(g("Kernel Name")[:60], "time", g("gpu__time_duration.sum"), units[hdr.index("gpu__time_duration.sum")],
      "regs", g("launch__registers_per_thread"), "grid", g("launch__grid_size"), "smem", g("launch__shared_mem_per_block_dynamic"))
t=num("gpu__time_duration.sum")
dr,dw=num("dram__bytes_read.sum"),num("dram__bytes_write.sum")
inst=float(g("smsp__inst_executed.sum"))
print(f"  dram {dr/reps/1e3:.0f}+{dw/reps/1e3:.0f} KB/rep-step  -> {(dr+dw)/t/1e12:.2f} TB/s;  inst/rep-step {inst/reps:.0f};  issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active')}%  fp64 {g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')}%  icc_hit {g('sm__icc_request_hit_rate.pct')}%  gcc_instr {g('gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed')}%  lsu_wf {g('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed')}%  L2hit {g('lts__t_sector_hit_rate.pct')}%")
out=[]
for i,h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
        try: out.append((float(r[i]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
        except: pass
print("  stalls/issue:", ", ".join(f"{h} {v:.2f}" for v,h in sorted(out,reverse=True)[:8]))
