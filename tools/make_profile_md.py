"""Regenerate profiles/r2_summary.md from the committed raw artifacts (launch lists, full ncu capture, bench lines).

    python tools/make_profile_md.py > profiles/r2_summary.md
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda *a: os.path.join(ROOT, "profiles", *a)      # noqa: E731

IN_STEP = ("prep_kernel", "tc_fwd_kernel", "fwd_finalize", "bwd_prep", "tc_bwd_ds", "bwd_fused_finalize")


def launch_table(path, step_us=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, order = {}, []
    for r in rows[1:]:
        if r[ki] not in agg:
            order.append(r[ki])
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    out = ["| kernel | launches in capture | mean duration (us) | " + (f"share of the graph-replayed step ({step_us:.0f} us) |" if step_us else "in the 6-launch step |"),
           "|---|---|---|---|"]
    for k in order:
        m = sum(agg[k]) / len(agg[k]) / 1000
        name = k.replace("void ", "").replace("tcelbo::", "")[:60]
        in_step = any(t in k for t in IN_STEP)
        share = (f"{100 * m / step_us:.2f} %" if step_us else "yes") if in_step else "no"
        out.append(f"| `{name}` | {len(agg[k])} | {m:.1f} | {share} |")
    return out


def main():
    j = json.load(open(P("r2_bench_n1.json")))
    step_us = j["ms_per_step"] * 1000
    km = j["roofline"]["kernel_ms"]
    out = ["# Round 2: launch lists, full ncu capture and bench lines of the shipped build (one B200, 1965 MHz)\n",
           "Workload: `bench.py` step = `compute_kl_loss`-equivalent evaluation (reparameterize + KL + TC(MSS) + mean + backward to mu / logvar) at "
           "global batch 8192, z_dim 128, N = 16 704; 6 library launches (`tcelbo_klloss_forward_ex` / `_backward_ex`).\n",
           "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -c 700`, `bench.py --steps 3 --warmup 3 --no-graph "
           "--no-cpu-baseline --no-train`; file `r2_launches.csv`)\n",
           "`--no-graph` issues the same C-ABI step the graph replays (`GraphedKLLoss(capture=False)`); the capture also contains bench.py's "
           "eager drop-in steps (`reparam_*`, the `at::` fill / add kernels of autograd), its L2 flush (`FillFunctor<unsigned char>`) and the ex2 probe.\n"]
    out += launch_table(P("r2_launches.csv"), step_us)
    out.append("")
    out.append(f"Live CUDA events inside `bench.py` on the same build (graph replay, `r2_bench_n1.json`): step {j['ms_per_step']:.3f} ms = "
               f"{j['value']:.4g} log-densities/s, forward sweep {km['tc_fwd_kernel']:.3f} ms ({100 * km['tc_fwd_kernel'] / j['ms_per_step']:.1f} %), "
               f"fused backward sweep {km['tc_bwd_ds_kernel']:.3f} ms ({100 * km['tc_bwd_ds_kernel'] / j['ms_per_step']:.1f} %): the ncu launch "
               f"list agrees on the shares.  Step as a fraction of the SFU roofline: {j['roofline']['step_frac_of_sfu_peak']:.3f}; drop-in path "
               f"(eager) {j['dropin']['ms_per_step']:.3f} ms; e2e with host buffers {j['e2e']['ms_per_step']:.3f} ms (graph) / "
               f"{j['e2e_dropin']['ms_per_step']:.3f} ms (drop-in); train (configs[1] shape) {j['train']['value']:.0f} images/s.\n")
    out.append("## Full capture (`ncu --set full --clock-control none --import-source on -k regex:tc_fwd_kernel|tc_bwd_ds_kernel "
               "--launch-skip 8 --launch-count 2`, same command; file `r2_ncu_full_raw.csv`)\n")
    out.append(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), P("r2_ncu_full_raw.csv")],
                              capture_output=True, text=True).stdout)
    out.append(open(P("r2_notes.md")).read())
    out.append("## Small batch (BASELINE configs[1] latent shape: B = 64, z_dim 128; `tools/small_batch_bench.py`, file `r2_launches_b64.csv`)\n")
    out += launch_table(P("r2_launches_b64.csv"))
    out.append("")
    out.append(open(P("r2_small_batch_notes.md")).read())
    out.append("## Bench lines (`bench.py --gpus N`, strong scaling at global batch 8192)\n")
    out.append("Train column: Soft-Intro-TC images/s of the bench's training leg -- BASELINE configs[1] shape (64x64, z 128, batch 64) at N = 1, "
               "configs[4] shape (128x128, z 256, batch 32 per GPU, data parallel) at N > 1, so the N = 1 entry is not the base of the others.\n")
    out.append("| N | ms / step | log-densities/s | efficiency vs N=1 | drop-in (eager) ms | e2e ms | e2e drop-in ms | train images/s | file |")
    out.append("|---|---|---|---|---|---|---|---|---|")
    base = None
    for n, f in SCALE_FILES:
        if not os.path.exists(P(f)):
            continue
        b = json.loads([l for l in open(P(f)) if l.startswith("{")][-1])
        base = base or b["ms_per_step"]
        tr = (b.get("train") or {}).get("value")
        out.append(f"| {n} | {b['ms_per_step']:.3f} | {b['value']:.4g} | {base / (n * b['ms_per_step']):.3f} | {b['dropin']['ms_per_step']:.3f} | "
                   f"{b['e2e']['ms_per_step']:.3f} | {b['e2e_dropin']['ms_per_step']:.3f} | {tr:.0f} | `{f}` |" if tr else
                   f"| {n} | {b['ms_per_step']:.3f} | {b['value']:.4g} | {base / (n * b['ms_per_step']):.3f} | {b['dropin']['ms_per_step']:.3f} | "
                   f"{b['e2e']['ms_per_step']:.3f} | {b['e2e_dropin']['ms_per_step']:.3f} | - | `{f}` |")
    print("\n".join(out))


SCALE_FILES = [(1, "r2_bench_n1.json"), (2, "r2_bench_n2.json"), (4, "r2_bench_n4.json"), (8, "r2_bench_n8.json")]


if __name__ == "__main__":
    main()
