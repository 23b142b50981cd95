"""Turn gpurun_out/configs_<tag>.json (bench.py --configs) into profiles/<tag>_configs.md."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
d = json.loads(open(os.path.join(ROOT, "gpurun_out", f"configs_{tag}.json")).read().strip().splitlines()[-1])
with open(os.path.join(ROOT, "profiles", f"{tag}_configs.md"), "w") as f:
    f.write(f"# The reference's own configurations on one B200 ({tag})\n\n`python bench.py --configs`: residual loss + backward to grad alpha through the "
            "reference-facing API (`train_api.py`, row-major `[B, N]` tensors as the reference passes them, B = 1000), CUDA events, 20 iterations; "
            f"beside it the oracle's CPU port (numpy/scipy fp32, {d['cores']} host cores visible) on the same inputs.  Differences are against the oracle in fp64 "
            "(tolerances: loss 1e-5, gradient 1e-4).  These are parity-test cases, not the bench line.\n\n"
            "| config | N | GPU ms (fwd+bwd) | GPU samples/s | CPU port ms | CPU port samples/s | ratio | loss rel. diff | grad rel. diff | path |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for r in d["configs"]:
        f.write(f"| {r['config']} | {r['N']} | {r['gpu_ms_fwd_bwd']:.3f} | {r['gpu_samples_per_s']:.3g} | {r['cpu_port_ms_fwd_bwd']:.1f} | "
                f"{r['cpu_port_samples_per_s']:.3g} | {r['speedup']:.0f} x | {r['loss_rel_diff_vs_fp64']:.1e} | {r['grad_rel_diff_vs_fp64']:.1e} | {r['note']} |\n")
    f.write(f"\nReference's own code (torch CPU, its Python loops): {d['reference_own_code']}.\n")
    L = d.get("linear_large")
    if L:
        f.write(f"\n{L['config']}: forward {L['fwd_ms']:.2f} ms ({L['fwd_algorithmic_GBs']:.0f} GB/s of 12 N B), backward {L['bwd_ms']:.2f} ms "
                f"({L['bwd_algorithmic_GBs']:.0f} GB/s of 8 N B; measured HBM peak {L['hbm_peak_GBs']:.0f} GB/s), {L['samples_per_s']:.0f} samples/s.\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_configs.md")).read())
