"""Summarises an .ncu-rep (raw + source pages) into the text kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
def page(name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))
raw = page("raw")
hdr, units = raw[0], raw[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum"]
for r in raw[2:]:
    d = dict(zip(hdr, r))
    print("== kernel", d.get("Kernel Name"), file=out)
    for k in KEYS:
        if k in d and d[k] not in ("", "n/a"):
            print(f"{k:75s} {d[k]:>18s} {units[hdr.index(k)]}", file=out)
src = page("source")
h = src[1]; ix = {n: i for i, n in enumerate(h)}
ops = collections.Counter(); tot = thr = samp = 0; stalls = collections.Counter()
scol = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
nstatic = 0
for r in src[2:]:
    try:
        n = int(r[ix["Instructions Executed"]]); t = int(r[ix["Thread Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except Exception:
        continue
    nstatic += 1
    w = r[ix["Source"]].split()
    o = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
    ops[o] += n; tot += n; thr += t; samp += s
    for c in scol:
        try: stalls[c] += int(r[ix[c]])
        except Exception: pass
print(f"\nstatic SASS instructions {nstatic}; executed warp instructions {tot}; average active threads per instruction {thr / max(tot, 1):.2f}", file=out)
print("instruction mix (executed):", file=out)
for o, n in ops.most_common(16):
    print(f"  {o:10s} {100 * n / tot:5.1f}%", file=out)
print("warp stall samples:", file=out)
for c, n in stalls.most_common(8):
    print(f"  {c:28s} {100 * n / max(samp, 1):5.1f}%", file=out)
