"""The oracle's "faithful" storage mode: every render target rounded to the texture format the reference allocates
it in (R8_UNORM AO, R16F / RGBA16F SSR + TAA, R11G11B10F Bloom, RG16F closest motion, RG8 blue noise). It is a SENSITIVITY figure - how far
the reference's own narrow targets move the frame - reported beside the kernel PSNR by bench.py; the parity gate stays the fp32 oracle.
"""
import numpy as np

from diligentfx_b200 import synth
from oracle import oracle_py as op


def test_quantisers_known_answers():
    # half: agrees with IEEE binary16 (numpy) incl. ties-to-even, denormals and overflow
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.standard_normal(2000).astype(np.float32) * s for s in (1e-6, 1e-3, 1.0, 1e3, 7e4)] +
                          [np.array([0.0, -0.0, 1.0, 65504.0, 65519.9, 65520.0, 6.1e-5, 5.96e-8, 2.98e-8, 2.99e-8, 1.0009765625, 1.00048828125], np.float32)])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).astype(np.float32)
    got = np.array([op.quantize(op.QUANT_HALF, float(v)) for v in vals], np.float32)
    assert np.array_equal(got, want)
    # UNORM8: round-to-nearest of saturate(v) * 255, NaN -> 0
    for v, w in ((0.0, 0.0), (1.0, 1.0), (2.0, 1.0), (-1.0, 0.0), (0.5 / 255.0 - 1e-6, 0.0), (0.5 / 255.0 + 1e-6, 1.0 / 255.0), (float("nan"), 0.0)):
        assert abs(op.quantize(op.QUANT_UNORM8, v) - w) < 1e-7, v
    # unsigned 11 / 10-bit floats: 6 / 5 mantissa bits, no sign, finite maximum
    assert op.quantize(op.QUANT_FLOAT11, 1.0) == 1.0 and op.quantize(op.QUANT_FLOAT10, 1.0) == 1.0
    assert op.quantize(op.QUANT_FLOAT11, 1.0 + 1.0 / 64) == 1.0 + 1.0 / 64 and op.quantize(op.QUANT_FLOAT10, 1.0 + 1.0 / 32) == 1.0 + 1.0 / 32
    assert op.quantize(op.QUANT_FLOAT11, 1.0 + 1.0 / 128) == 1.0            # tie -> even mantissa
    assert op.quantize(op.QUANT_FLOAT11, -3.0) == 0.0 and op.quantize(op.QUANT_FLOAT10, -1e-3) == 0.0
    assert op.quantize(op.QUANT_FLOAT11, 1e9) == 65024.0 and op.quantize(op.QUANT_FLOAT10, 1e9) == 64512.0
    for f in (op.QUANT_HALF, op.QUANT_FLOAT11, op.QUANT_FLOAT10, op.QUANT_UNORM8, op.QUANT_UNORM16):  # idempotent
        for v in (0.1234, 3.75, 1e-4, 0.999):
            q = op.quantize(f, v)
            assert op.quantize(f, q) == q


def test_faithful_chain_sensitivity():
    w, h = 192, 108
    seq = synth.generate_sequence(w, h, 4)
    a, b = op.Oracle(w, h, threads=4), op.Oracle(w, h, threads=4)
    b.set_storage(True)
    for fr in seq:
        for o in (a, b):
            o.set_inputs(fr)
            o.frame()
    # stored planes hold representable values only
    ao = b.get("ssao_out")
    assert np.array_equal(np.round(ao * 255.0) / 255.0, ao.astype(np.float64).astype(np.float32)) or np.abs(np.round(ao * 255.0) / 255.0 - ao).max() < 1e-7
    acc = b.get("taa_accum1")
    assert np.array_equal(acc, acc.astype(np.float16).astype(np.float32))
    # and the frame moves by a visible but small amount: this is the figure bench.py prints as psnr.reference_storage_vs_fp32
    la, lb = np.clip(a.get("ldr")[..., :3], 0, 1).astype(np.float64), np.clip(b.get("ldr")[..., :3], 0, 1).astype(np.float64)
    mse = float(np.mean((la - lb) ** 2))
    p = 10.0 * np.log10(1.0 / mse)
    print(f"LDR PSNR, reference render-target formats vs fp32 planes, {w}x{h} frame 4: {p:.2f} dB")
    assert 25.0 < p < 90.0
    assert not np.array_equal(a.get("ssao_out"), ao)
