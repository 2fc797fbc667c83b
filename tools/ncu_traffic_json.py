"""profiles/r2d_ncu_raw.csv (ncu -i ... --page raw --csv of the `--set full` capture) -> profiles/r2d_ncu_traffic.json: DRAM bytes per launch of each
kernel (steady-state launch = the first of each kernel in the capture), the figure bench.py reports as roofline.traffic."""
import csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2d_ncu_raw.csv")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2d_ncu_traffic.json")
rd = csv.reader(open(src))
hdr, units = next(rd), next(rd)
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1, "us": 1e3, "ms": 1e6}
kernels = {}
for r in rd:
    name = re.sub(r"^void ", "", r[col["Kernel Name"]])
    name = name[:name.index(">(") + 1] if ">(" in name else name[:name.index("(")]
    name = name.replace("(int)", "")
    if name in kernels:
        kernels[name]["launches_profiled"] += 1
        continue
    rdb = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
    wrb = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
    kernels[name] = {"dram_bytes": int(rdb + wrb), "dram_read_bytes": int(rdb), "dram_write_bytes": int(wrb),
                     "duration_ns": int(float(r[col["gpu__time_duration.sum"]]) * tscale[units[col["gpu__time_duration.sum"]]]), "launches_profiled": 1}
pass_to_kernel = {
    "ssr_intersect": "ssr_intersect_kernel<0, 0, 0, 1>", "ssr_bilateral": "ssr_bilateral_kernel<1>", "ssao_ambient_occlusion": "ssao_ao_kernel<0, 1>",
    "ssao_temporal": "ssao_temporal_kernel", "ssao_resample": "ssao_resample_kernel<1>", "ssao_spatial": "ssao_spatial_tile_kernel<1>",
    "compose_taa": "taa_kernel<1, 0, 0, 1, 1>", "bloom_prefilter": "bloom_down2x_stream_kernel<1>", "bloom_composite_tonemap": "bloom_up2x_stream_kernel<1, 1>",
}
json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch at 3840x2160, G-buffer in the renderer formats, from `ncu --set full --clock-control none` "
                       "(profiles/r2d_ncu_raw.csv, summarised in profiles/r2d_ncu_full.md); first (steady-state) launch of each kernel. bench.py reads this file for roofline.traffic.",
           "kernels": kernels, "pass_to_kernel": {k: v for k, v in pass_to_kernel.items() if v in kernels}}, open(dst, "w"), indent=1)
print(json.dumps({k: v["dram_bytes"] for k, v in kernels.items()}, indent=1))
