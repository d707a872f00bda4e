"""profiles/r02_scale_c5.md from the per-N bench lines (gpurun_out/bench_n{1,2,4,8}.json; N=1 may be the plain bench line).
    python tools/scale_report.py gpurun_out/bench9.json gpurun_out/bench_n2.json ..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit("no JSON line in " + path)


rows = sorted((load(p) for p in sys.argv[1:]), key=lambda d: d["n_gpus"])
base = rows[0]
out = ["# Round 2 -- multi-GPU: C2 headline (independent scenes per rank) and the C5 data path with its one collective", "",
       "One process per GPU (`torch.distributed.run`, NCCL), `bench.py --gpus N --steps 50`; device times are CUDA events, max over ranks.",
       "", "## C2 headline (`value`: no data-path collective, weak scaling) and end-to-end", "",
       "| N | value views/s | per GPU | efficiency vs N=%d | e2e views/s | e2e per GPU | e2e ms/step median |" % base["n_gpus"],
       "|---:|---:|---:|---:|---:|---:|---:|"]
for d in rows:
    n = d["n_gpus"]
    out.append("| %d | %.0f | %.0f | %.3f | %.0f | %.0f | %.3f |" % (
        n, d["value"], d["value"] / n, (d["value"] / n) / (base["value"] / base["n_gpus"]), d["e2e"]["value"], d["e2e"]["value"] / n,
        d["e2e"].get("ms_per_step_median", float("nan"))))
out += ["", "## C5 data path: per rank 1 sample -> VAE decode -> ONE ncclAllGather of the decoded surfels -> sharded render of N x 8 views of 512^2", "",
        "| N | views/s | per GPU | efficiency | decode ms | all-gather us (max over ranks) | bytes gathered | algbw GB/s | render ms | cascade samples/s (2x249 NFE + decode + gather + render) | per GPU |",
        "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
b5 = base["c5"]
for d in rows:
    c, n = d["c5"], d["n_gpus"]
    st, co = c["stage_ms_rank0"], c["collective"]
    cas = c.get("cascade", {}).get("reference_cfg_2B_both_stages", {})
    out.append("| %d | %.0f | %.0f | %.3f | %.2f | %.1f | %.2f MB | %s | %.2f | %s | %s |" % (
        n, c["views_per_s"], c["views_per_s"] / n, (c["views_per_s"] / n) / (b5["views_per_s"] / base["n_gpus"]), st["vae_decode"],
        co["us_max_over_ranks"], co["bytes_total"] / 1e6, ("%.0f" % co["algbw_GBs"]) if co.get("algbw_GBs") else "-", st["render_shard"],
        ("%.3f" % cas["samples_per_s"]) if cas else "-", ("%.3f" % (cas["samples_per_s"] / n)) if cas else "-"))
out += ["", "The all-gather moves 3.83 MB per rank.  Its time is the interval between two CUDA events on the compute stream around",
        "torch.distributed's call (NCCL's own stream, ordered against the compute stream by events), so it contains the stream",
        "hand-off AND the wait for the slowest rank's decode to reach the collective (arrival skew), not only the NVLink transfer",
        "(11.5 MB received per rank at N = 4 is ~15 us at the measured 770 GB/s).  Rendering after the gather is rank-local: each",
        "rank renders 8 of the N x 8 (sample, view) pairs in one batched launch set.", ""]
path = os.path.join(ROOT, "profiles", "r02_scale_c5.md")
open(path, "w").write("\n".join(out))
print(path)
