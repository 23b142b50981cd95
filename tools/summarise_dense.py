"""Turn the outputs of tools/profile_dense.sh / tools/dense_sweep.sh (gpurun_out/) into profiles/<tag>_dense_ncu_summary.md."""
import csv, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rows = list(csv.reader(open(os.path.join(G, f"prof_dense_{tag}_raw.csv"))))
H, U, V = rows[0], rows[1], rows[2]
val = {h: (v, u) for h, u, v in zip(H, U, V)}
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum"]
kernel = val.get("Kernel Name", ("?", ""))[0]
out = [f"# ncu --set full summary of the preconditioner GEMM ({tag})", "",
       "`ncu --set full --clock-control none --import-source on -k regex:dense_apply_tc -s 6 -c 1` on `python tools/time_dense.py 2549 1024 20` "
       "(cfg2b size: N = 2549, B = 1024, fused `- F` and loss epilogue); `tools/profile_dense.sh`, summarised by `tools/summarise_dense.py`.  "
       "Times under ncu are not bench values.", "", f"Kernel: `{kernel}`", "", "| metric | value | unit |", "|---|---|---|"]
for k in want:
    if k in val:
        out.append(f"| `{k}` | {val[k][0]} | {val[k][1]} |")
out += ["", "Launch list of one apply (`ncu --metrics gpu__time_duration.sum`, cold-cache and serialised: shares, not absolutes):", "", "```"]
lr = [r for r in csv.reader(open(os.path.join(G, f"dense_launches_{tag}.csv"))) if r and not r[0].startswith("==")]
LH = lr[0]
for r in lr[1:4]:
    d = dict(zip(LH, r))
    out.append(f"{float(d['Metric Value']) / 1e3:8.1f} us  {d['Kernel Name'][:90]}  grid {d['Grid Size']}")
out += ["```", "", "Plain timing of the same command (CUDA events, no profiler; second generation, first generation `FEO_DENSE_GEN=1`, fp32 FMA kernel `FEO_DENSE_SIMT=1`):", "", "```"]
out += [l.rstrip() for l in open(os.path.join(G, f"dense_plain_{tag}.log")) if l.strip()]
out += ["```"]
for mode, title in (("gens", "Kernel generations (default = launcher's choice, pairs, pre-split single CTAs, first generation, fp32 FMA) and other sizes, pairs vs single CTAs"),
                    ("tiles", "Second generation: tile widths, stages, cluster multicast, refill gap, drains, developer timing modes"),
                    ("pairs", "Third generation (CTA pairs): k-blocks per stage, stages, drains, clusters of two pairs, 192-column tiles, no-MMA floor"),
                    ("small", "Small sizes: launch-bound, all variants within the noise of a three-kernel Python loop")):
    p = os.path.join(G, f"dense_sweeps_{mode}.log")
    if os.path.exists(p):
        out += ["", f"{title} (`tools/dense_sweeps.sh {mode}`):", "", "```"] + [l.rstrip()[:175] for l in open(p) if l.strip()] + ["```"]
extra = os.path.join(P, f"{tag}_dense_notes.md")
if os.path.exists(extra):
    out += ["", open(extra).read().rstrip()]
open(os.path.join(P, f"{tag}_dense_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("wrote", os.path.join(P, f"{tag}_dense_ncu_summary.md"))
