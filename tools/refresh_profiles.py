"""Copy one GPU visit's outputs (gpurun_out/<tag>_*) into profiles/r02_* and rebuild profiles/r02_ncu_capture_summary.json, which bench.py
reads for the roofline.limiter / *_pct fields.  usage: python tools/refresh_profiles.py <tag>"""
import json, re, shutil, subprocess, sys

tag = sys.argv[1]
src, dst = "gpurun_out", "profiles"
for wl in ["headline", "C2"]:
    for k in ["icp_fused_kernel", "lcp_score_kernel"]:
        shutil.copy(f"{src}/{tag}_ncu_{k}_{wl}.txt", f"{dst}/r02_ncu_{k}_{wl}.txt")
    shutil.copy(f"{src}/{tag}_launches_{wl}.csv", f"{dst}/r02_launches_{wl}.csv")
    with open(f"{dst}/r02_launches_{wl}.txt", "w") as f:
        subprocess.run([sys.executable, "tools/launch_summary.py", f"{dst}/r02_launches_{wl}.csv"], stdout=f, stderr=subprocess.STDOUT)
for a, b in [("bench_default.json", "bench_default.json"), ("bench_reference.json", "bench_reference.json"), ("pytest_gpu.log", "pytest_gpu.log"), ("smoke.log", "smoke.log")]:
    shutil.copy(f"{src}/{tag}_{a}", f"{dst}/r02_{b}")


def parse(p):
    d = {}
    for l in open(p):
        m = re.match(r"\s+(\S+)\s+([0-9.]+)\s*(\S*)", l)
        if m:
            d[m.group(1)] = (float(m.group(2)), m.group(3))
    return d


out = {}
for wl in ["headline", "C2"]:
    d = parse(f"{dst}/r02_ncu_icp_fused_kernel_{wl}.txt")
    b = lambda k: d[k][0] * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d[k][1]]
    t, u = d["gpu__time_duration.sum"]
    l1, iss = d["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"][0], d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]
    lim = "l1tex" if l1 >= 60 else "latency (the LM chains of the slowest hypotheses; l1tex %.0f %%, issue %.0f %%)" % (l1, iss)
    out[wl] = {"dram_bytes_per_launch": int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum")), "limiter": lim, "l1tex_pct": l1,
               "dram_pct": d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0], "warps_active_pct": d["sm__warps_active.avg.pct_of_peak_sustained_active"][0],
               "issue_active_pct": iss, "l1tex_hit_pct": d["l1tex__t_sector_hit_rate.pct"][0], "lts_pct": d["lts__throughput.avg.pct_of_peak_sustained_elapsed"][0],
               "kernel_ms_under_ncu": t * {"ms": 1, "us": 1e-3, "s": 1e3}[u],
               "source": f"profiles/r02_ncu_icp_fused_kernel_{wl}.txt (ncu --set full --clock-control none, one launch after warm-up)"}
json.dump(out, open(f"{dst}/r02_ncu_capture_summary.json", "w"), indent=1)
print(json.dumps(out, indent=1))
