"""Markdown tables of DESIGN.md §4 from a bench line (profiles/<file>.json) and the ncu summary numbers (profiles/r2d_ncu_full.md).
usage: python tools/design_tables.py profiles/r2j_bench_n1.json"""
import json, sys

r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
W, H = r["config"]["width"], r["config"]["height"]
NCU = {  # pass -> (kernel, instr/px steady, issue % steady, instr/px live, issue % live)   [profiles/r2d_ncu_full.md]
    "ssr_intersect": ("ssr_intersect_kernel", 1871, 78, None, None),
    "ssr_bilateral": ("ssr_bilateral_kernel", 285, 54, None, None),
    "ssao_ambient_occlusion": ("ssao_ao_kernel", 1175, 81, None, None),
    "ssao_temporal": ("ssao_temporal_kernel", 412, 75, None, None),
    "ssao_resample": ("ssao_resample_kernel", 85, 25, 477, 71),
    "ssao_spatial": ("ssao_spatial_tile_kernel (TMA)", 360, 76, 451, 81),
    "compose_taa": ("taa_kernel<COMPOSE>", 857, 74, None, None),
    "bloom_prefilter": ("bloom_down2x_stream_kernel<1>", 65, 56, None, None),
    "bloom_composite_tonemap": ("bloom_up2x_stream_kernel<1, TM>", 182, 79, None, None),   # profiles/r2j_bloom_levels.md (final launch shape)
}
KERNEL = {
    "blue_noise": "blue_noise_kernel (128x128)", "postfx_prepare": "postfx_prepare_kernel (P1-P3 fused)", "ssr_hiz": "pyramid_tile_kernel<HizOp> (TMA) + cluster tail",
    "ssr_mask_roughness": "ssr_mask_kernel", "ssr_spatial": "ssr_spatial_kernel", "ssr_temporal": "ssr_temporal_kernel",
    "ssao_prefilter_depth": "pyramid_tile_kernel<PrefilterOp> (TMA)", "ssao_convolute": "pyramid_tile_kernel<ConvoluteOp> (TMA)",
    "bloom_downsample": "bloom_down2x_stream_kernel<0> (levels 1-3), bloom_downsample_kernel (4-5): 5 launches", "bloom_tail": "bloom_tail_kernel (cluster)",
    "bloom_upsample": "bloom_upsample_kernel (levels 4-3), bloom_up2x_stream_kernel<0> (2-0): 5 launches",
}
print("| pass | kernel | B/px | ms | % HBM peak | live: ms (% peak) | instr/px | issue % | bound |")
print("|---|---|---:|---:|---:|---:|---:|---:|---|")
for p in r["passes"]:
    n = NCU.get(p["pass"])
    kern = n[0] if n else KERNEL.get(p["pass"], "")
    live = p.get("live")
    live_s = "%.4f (%.0f)" % (live["ms"], 100 * live["frac"]) if live and abs(live["ms"] - p["ms"]) > 0.1 * p["ms"] else "="
    ins = ("%d" % n[1] + (" / %d live" % n[3] if n[3] else "")) if n else ""
    iss = ("%d" % n[2] + (" / %d" % n[4] if n[4] else "")) if n else ""
    if n:
        bound = ("HBM + issue" if p["frac"] >= 0.5 else "issue") if n[2] >= 70 or (n[4] or 0) >= 70 else ("HBM" if p["frac"] >= 0.5 else "latency / occupancy")
    else:
        bound = "HBM" if p["frac"] >= 0.5 else ("launch + L2 latency (small planes)" if p["alg_bytes"] / (W * H) < 13 else "latency / occupancy")
    print("| %s | `%s` | %.1f | %.4f | %.0f | %s | %s | %s | %s |" % (p["pass"], kern, p["alg_bytes"] / (W * H), p["ms"], 100 * p["frac"], live_s, ins, iss, bound))
tot = sum(p["ms"] for p in r["passes"])
print("| **sum of passes (serial)** | | %.1f | %.4f | %.0f | | | | |" % (sum(p["alg_bytes"] for p in r["passes"]) / (W * H), tot,
      100 * sum(p["alg_bytes"] for p in r["passes"]) / (tot * 1e-3) / 1e9 / r["roofline"]["peak"]))
print("| **frame, three streams + graphs (`value`)** | | | %.4f | %.0f | | | | |" % (r["ms_per_step"], 100 * r["roofline"]["chain"]["frac"]))
