"""profiles/r2_roofline_profile.json from an ncu --set full capture of the shipped kernels (bench.py's roofline.traffic
and issue fractions come from it; bench.py refuses it when the kernel sources have changed since).

On the GPU box (gpurun):
    ncu --set full --clock-control none -k regex:"preprocess_|color_fwd|scan_tiles|scatter_kernel|tile_sort|render_|long_tile|onesweep" \
        -s 100 -c 20 -o gpurun_out/r2_profile python tests/tools/stage_probe.py
Here:
    python tests/tools/make_roofline_profile.py gpurun_out/r2_profile.ncu-rep
"""
import csv, io, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
ix = {n: i for i, n in enumerate(h)}
units = rows[1]


def val(row, name):
    v = float(row[ix[name]].replace(",", ""))
    u = units[ix[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "second": 1e6}.get(u, 1.0)
    return v * scale


agg = {}
for row in rows[2:]:
    name = row[ix["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").strip()
    a = agg.setdefault(short, dict(n=0, us=0.0, inst=0.0, rd=0.0, wr=0.0, regs=0, issue=0.0, warps=0.0))
    a["n"] += 1
    a["us"] += val(row, "gpu__time_duration.sum")
    a["inst"] += val(row, "smsp__inst_executed.sum")
    a["rd"] += val(row, "dram__bytes_read.sum")
    a["wr"] += val(row, "dram__bytes_write.sum")
    a["regs"] = int(float(row[ix["launch__registers_per_thread"]]))
    a["issue"] += float(row[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
    a["warps"] += float(row[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]])
kernels = [dict(name=k, launches=a["n"], us_per_launch=round(a["us"] / a["n"], 3), warp_instructions_per_launch=round(a["inst"] / a["n"]),
                dram_read_bytes_per_launch=round(a["rd"] / a["n"]), dram_write_bytes_per_launch=round(a["wr"] / a["n"]),
                dram_bytes_per_launch=round((a["rd"] + a["wr"]) / a["n"]), registers=a["regs"],
                issue_active_pct=round(a["issue"] / a["n"], 2), warps_active_pct=round(a["warps"] / a["n"], 2)) for k, a in sorted(agg.items())]
git = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
dirty = subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", *["gs_localization_b200/csrc/" + f for f in bench.PROFILED_SOURCES], "include"],
                       capture_output=True, text=True).stdout.strip()
doc = {"_comment": "per-launch averages from one ncu --set full --clock-control none capture of tests/tools/stage_probe.py (headline workload); "
                   "times are under the profiler (cold caches, serialised): use the byte and instruction counts, not the times",
       "source_sha256": bench.kernel_source_sha256(), "git": git + ("+uncommitted kernel changes" if dirty else ""),
       "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "report": os.path.basename(rep), "kernels": kernels}
json.dump(doc, open(bench.PROFILE_JSON, "w"), indent=1)
print(json.dumps(doc, indent=1)[:3000])
