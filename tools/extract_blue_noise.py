#!/usr/bin/env python3
"""Extract the Heitz et al. blue-noise sampler tables into a binary data blob.

The tables (Owen-scrambled Sobol sequence `Sobol_256d[256]` and the optimised
`ScramblingTile[128*128*8]`, both uint8) are *data* that the reference uploads into
two R8_UINT textures (reference: PostProcess/Common/src/PostFXContext.cpp:151-191,
table file PostProcess/Common/src/SamplerBlueNoiseErrorDistribution_128x128_OptimizedFor_2d2d2d2d_1spp.cpp).
They are the result of an offline optimisation (Heitz, Belcour, Ostromoukhov, Coeurjolly,
Iehl: "A Low-Discrepancy Sampler that Distributes Monte Carlo Errors as a Blue Noise in
Screen Space", SIGGRAPH Talks 2019) and cannot be regenerated algorithmically, so the
numbers themselves are shipped as a 131,328-byte blob:

    bytes [0, 256)          Sobol_256d
    bytes [256, 256+131072) ScramblingTile

Run in the build container (the reference tree is not available on the GPU box):
    python tools/extract_blue_noise.py /root/reference diligentfx_b200/data/blue_noise_tables.bin
"""
import re
import sys
import hashlib


def main(ref_root: str, out_path: str) -> None:
    src = (ref_root.rstrip("/") +
           "/PostProcess/Common/src/SamplerBlueNoiseErrorDistribution_128x128_OptimizedFor_2d2d2d2d_1spp.cpp")
    text = open(src, "r").read()

    def table(name: str, count: int) -> bytes:
        m = re.search(name + r"\s*\[[^\]]*\]\s*=\s*\{([^}]*)\}", text)
        if m is None:
            raise SystemExit(f"table {name} not found in {src}")
        vals = [int(v) for v in m.group(1).replace("\n", " ").split(",") if v.strip()]
        if len(vals) != count:
            raise SystemExit(f"table {name}: expected {count} entries, found {len(vals)}")
        return bytes(vals)

    blob = table("Sobol_256d", 256) + table("ScramblingTile", 128 * 128 * 8)
    with open(out_path, "wb") as f:
        f.write(blob)
    print(out_path, len(blob), "bytes sha256", hashlib.sha256(blob).hexdigest())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
