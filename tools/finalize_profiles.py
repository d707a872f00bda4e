"""Turns the outputs of profiles/run_r02.sh (gpurun_out/r02_*) into the tracked files under profiles/.
    python tools/finalize_profiles.py"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def first_json(path):
    for line in open(path):
        if line.strip().startswith("{"):
            return line.strip()
    return None


def summarize(launches, rep, out, title):
    if os.path.exists(os.path.join(G, launches)) and os.path.exists(os.path.join(G, rep)):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), os.path.join(G, launches),
                        os.path.join(G, rep), os.path.join(P, out), title], check=False)
    else:
        print("skip", out)


for src, dst in (("r02_bench.json", "r02_bench_final.json"), ("r02_bench_reference.json", "r02_bench_reference.json")):
    p = os.path.join(G, src)
    if os.path.exists(p) and first_json(p):
        open(os.path.join(P, dst), "w").write(first_json(p) + "\n")
for f in ("r02_pytest_gpu.txt", "r02_smoke.txt", "r02_gpu.csv"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
summarize("r02_launches_raster.csv", "r02_render.ncu-rep", "r02_raster.md",
          "Round 2 - surfel rasteriser, C2 workload (tools/raster_variants.py), one launch = 6 views, list_k = 32")
summarize("r02_launches_dit.csv", "r02_attn.ncu-rep", "r02_attention.md", "Round 2 - DiT launch list (C3 + deployed legs) and the attention kernel")
summarize("r02_launches_dit.csv", "r02_gemm_mlp1.ncu-rep", "r02_gemm.md", "Round 2 - tcgen05 GEMM 4096x3072x768 + GELU")
summarize("r02_launches_n1.csv", "r02_micro_attn.ncu-rep", "r02_n1.md", "Round 2 - row N1 (VAE decoder) launch list and the micro-attention kernel")
# DRAM traffic per launch of the raster kernels (bench.py's roofline.traffic)
rep = os.path.join(G, "r02_render.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    tr = {"_source": "profiles/r02_raster.md (ncu --set full, tools/raster_variants.py C2 workload, one launch = 6 views)"}

    def mb(r, k):
        v = float(r[h.index(k)].replace(",", ""))
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u[h.index(k)], 1.0)
    for r in rr[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
        tr.setdefault(name, {"dram_bytes": 0.0, "launches": 0})
        tr[name]["dram_bytes"] += mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum")
        tr[name]["launches"] += 1
    bwd = sum(v["dram_bytes"] for k, v in tr.items() if k.startswith("render_bwd"))
    fwd = sum(v["dram_bytes"] for k, v in tr.items() if k.startswith("render_fwd"))
    tr["render_bwd_kernel"] = {"dram_bytes": bwd, "note": "sum over the backward's kernels of one step (A list-walking, A recompute, B x2, fused no-op)"}
    tr["render_fwd_kernel"] = {"dram_bytes": fwd}
    json.dump(tr, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print("traffic.json", {k: v.get("dram_bytes") for k, v in tr.items() if isinstance(v, dict)})
