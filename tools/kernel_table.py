"""profiles/rNN_kernel_table.md from a bench.py line (other_ops, legacy_gpu, roofline, networks).
    python tools/kernel_table.py profiles/r02_bench_n1.json > profiles/r02_kernel_table.md"""
import json
import sys

d = json.load(open(sys.argv[1]))
leg = {e["op"].replace(" (incl. zero fills)", ""): e for e in d.get("legacy_gpu", [])}
out = ["# Kernel table (one B200, `python bench.py`, CUDA events, back-to-back launches on inputs larger than L2)", "",
       "Source: `%s` (`other_ops`, `legacy_gpu`, `roofline`, `roofline_fwd`).  `frac` = algorithmic bytes / time / %.1f GB/s" % (sys.argv[1], d["roofline"]["peak"]),
       "(`MEASURED_PEAKS.json`).  legacy = the reference's `my_lib_kernel.cu` recompiled for sm_100a (the `legacy_gpu` leg of bench.py), same inputs and",
       "harness, including the zero fills its contract needs.", "",
       "| op | B/px | ours ms | ours frac | legacy ms | speed-up |", "|---|---|---|---|---|---|"]


def row(op, bpp, ms, frac):
    e = leg.get(op)
    out.append("| %s | %d | %.3f | %.3f | %s | %s |" % (op, bpp, ms, frac, "%.3f" % e["legacy_ms"] if e else "-",
                                                         "%.2fx" % (e["legacy_ms"] / ms) if e else "-"))


rf, rb = d["roofline_fwd"], d["roofline"]
row("FilterInterpolation forward 1920x1080, C=3, batch 4", 96, rf["ms_per_launch"], rf["frac"])
row("FilterInterpolation backward 1920x1080, C=3, batch 4", 180, rb["ms_per_launch"], rb["frac"])
for e in d.get("other_ops", []):
    row(e["op"], e["alg_bytes_per_px"], e["ms"], e["frac"])
out += ["", "Headline step (fwd + zero fill of gradinput1 + bwd): %.3f ms = %.0f Mpx/s; end to end through pinned host buffers: %.0f Mpx/s"
        % (d["ms_per_step"], d["value"], d["e2e"]["value"]),
        "(%.1f GB/s each way over PCIe); reference CPU implementation on %d host threads: %.1f Mpx/s."
        % (d["e2e"]["h2d_gbs_per_rank"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["value"])]
if d.get("networks"):
    out += ["", "Networks (random init, inference, one GPU; the same network object on this my_package and on the reference kernels):", "",
            "| network | frame | frame pairs | frames/s ours | frames/s reference kernels | max-abs / PSNR vs reference kernels |", "|---|---|---|---|---|---|"]
    for n in d["networks"]:
        out.append("| %s | %s | %d | %.2f | %.2f | %.3g / %.1f dB |" % (n["network"], n["frame"], n["frame_pairs_per_gpu"], n["frames_per_s"],
                                                                       n["frames_per_s_reference_kernels"], n["max_abs_vs_reference_kernels"],
                                                                       n["psnr_db_vs_reference_kernels"]))
print("\n".join(out))
