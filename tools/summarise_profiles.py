"""Turn the ncu outputs of tools/profile_bench.sh (gpurun_out/) into the tracked summaries under profiles/."""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def short_name(k):
    m = re.search(r"residual_lattice_kernel<\(?(?:bool\))?\s*(\w+)", k)
    if m:
        return "residual_lattice_kernel<" + ("bwd" if m.group(1) in ("1", "true") else "fwd") + ">"
    m = re.search(r"(residual_\w+|\w+_kernel)", k)
    return m.group(1) if m else k.split("(")[0].replace("void ", "")[:70]

G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list ------------------------------------------------------------------------------------
rows = list(csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
launches = [(r[ki], float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0)) for r in rows[h + 1:] if len(r) == len(H) and r[0].isdigit()]
tot = {}
for k, us in launches:
    name = short_name(k)
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += us
total_us = sum(v[1] for v in tot.values())
with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
    f.write(f"# ncu launch list ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 60` on "
            "`python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train-step --no-parity --no-precond-gemm` (cfg5, N=1 001 334, B=1024, one B200).\n"
            "Per-launch times are serialised and cold-cache: compare shares, not absolutes.\n\n"
            "| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{name}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total_us:.1f} % |\n")
    f.write("\nFirst 24 launches in order:\n\n```\n")
    for k, us in launches[:24]:
        f.write(f"{us:10.1f} us  {k[:100]}\n")
    f.write("```\n")

# ---- full capture -----------------------------------------------------------------------------------
rep = os.path.join(G, f"prof_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
traffic = {}
with open(os.path.join(P, f"{tag}_ncu_summary.md"), "w") as f:
    f.write(f"# ncu --set full summary ({tag})\n\n`ncu --set full --clock-control none --import-source on -k regex:residual_ -s 6 -c 2` on the same bench "
            "command; one launch of each fused kernel.  Times under ncu are not bench values.\n\n")
    for r in rr[2:]:
        d = dict(zip(hdr, r))
        name = short_name(d["Kernel Name"])
        f.write(f"## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
        for w in want[1:]:
            if w in d:
                f.write(f"| `{w}` | {d[w]} | {units[hdr.index(w)]} |\n")
        def gb(key):
            v = float(d[key].replace(",", "")); u = units[hdr.index(key)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
        traffic[name] = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        f.write(f"\nDRAM traffic per launch: {traffic[name] / 1e9:.2f} GB (algorithmic 12.30 GB).\n\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/_src.csv", "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), "/tmp/_src.csv", "12"], capture_output=True, text=True).stdout
    f.write("## Stall samples by SASS instruction (top 12 per kernel)\n\n```\n" + hot + "```\n")
traffic_named = dict(traffic)
traffic_named["source"] = f"ncu --set full --clock-control none, one launch each (tools/profile_bench.sh {tag}): dram__bytes_read.sum + dram__bytes_write.sum"
traffic_named["algorithmic_bytes_per_launch"] = 12304392192
json.dump(traffic_named, open(os.path.join(P, f"traffic_{tag}.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launches.md")).read()[:1500])
print(json.dumps(traffic))
