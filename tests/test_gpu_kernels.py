"""GPU: kernel-level checks through the C ABI (sdvg_gemm) against a float64 torch reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def ref(A, W, bias, relu, precision):
    if precision in ("fp16", "bf16"):
        dt = torch.float16 if precision == "fp16" else torch.bfloat16
        A, W = A.to(dt), W.to(dt)          # the kernel rounds operands to 16 bit, accumulates in fp32
    y = A.double() @ W.double().t()
    if bias is not None:
        y = y + bias.double()
    return torch.relu(y) if relu else y


# tolerance: fp32 accumulation error of a K-long dot product, relative to the output's max
TOL = {"fp32_simt": 5e-6, "fp32": 1e-5, "fp16": 1e-5, "bf16": 1e-5}


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "fp16", "bf16"])
@pytest.mark.parametrize("shape", [(128, 256, 64), (77, 96, 32), (300, 520, 2048), (1, 32, 8), (640, 6144, 256),
                                   (2000, 2048, 2048)])
def test_gemm_matches_float64(precision, shape):
    import sdvg_b200
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g)
    for relu in (False, True):
        C, _ = sdvg_b200.gemm(A, W, b, relu=relu, precision=precision)
        assert relerr(C, ref(A, W, b, relu, precision)) < TOL[precision]


@pytest.mark.parametrize("precision,bns", [("fp16", (32, 64, 128, 256, -64, -128, -192, -256)),
                                           ("fp32", (32, 64, 128, -64, -128))])
def test_gemm_every_tile_width_is_bit_identical(precision, bns):
    """The accumulation order along K does not depend on the tile shape: every instantiation of the one-CTA
    kernel (block_n > 0) and of the CTA-pair kernel (block_n < 0) agrees bit for bit."""
    import sdvg_b200
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(333, 1024, device="cuda", generator=g)
    W = torch.randn(512, 1024, device="cuda", generator=g) * 0.05
    outs = [sdvg_b200.gemm(A, W, None, precision=precision, block_n=bn)[0] for bn in bns]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_split_gemm_recovers_fp32_operands():
    """fp16 rounding of the operands costs ~5e-4; the split mode must be >100x closer to the exact product."""
    import sdvg_b200
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(256, 2048, device="cuda", generator=g) * 30.0     # un-normalised emb*sqrt(d) scale
    W = torch.randn(512, 2048, device="cuda", generator=g) * 0.02
    exact = A.double() @ W.double().t()
    e16 = relerr(sdvg_b200.gemm(A, W, None, precision="fp16")[0], exact)
    e32 = relerr(sdvg_b200.gemm(A, W, None, precision="fp32")[0], exact)
    assert e32 < 5e-6 and e16 > 50 * e32


def test_gemm_linearity():
    import sdvg_b200
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(130, 512, device="cuda", generator=g)
    W = torch.randn(256, 512, device="cuda", generator=g) * 0.05
    c1 = sdvg_b200.gemm(A, W, None, precision="fp32")[0]
    c2 = sdvg_b200.gemm(2.0 * A, W, None, precision="fp32")[0]      # power-of-two scaling is exact in every plane
    assert torch.equal(c2, 2.0 * c1)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_split_k_small_m_gemm(precision):
    """M <= 128 GEMMs split K over 2/4/8 CTAs per tile (partial tiles exchanged through L2, summed in a fixed
    order): same result as the float64 reference, deterministic from launch to launch, flags re-armed between
    launches (iters > 1 reuses them), ragged M / N, bias + ReLU in the fused epilogue of rank 0."""
    import sdvg_b200
    for (M, N, K) in ((96, 1024, 1024), (40, 1024, 2048), (128, 512, 1024), (80, 3072, 1024), (5, 200, 512)):
        g = torch.Generator(device="cuda").manual_seed(M + N)
        A = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g) * 0.05
        b = torch.randn(N, device="cuda", generator=g)
        want = ref(A, W, b, True, precision)
        for bn, ks in ((32, 4), (32, 2), (64, 2), (32, 8), (128, 4), (64, 4)):
            tiles = -(-N // bn)
            if tiles * ks > 148 or (K // 64) // ks < 1:
                continue
            C1, _ = sdvg_b200.gemm(A, W, b, relu=True, precision=precision, block_n=1000 * ks + bn)
            C2, _ = sdvg_b200.gemm(A, W, b, relu=True, precision=precision, block_n=1000 * ks + bn, iters=3)
            assert relerr(C1, want) < TOL[precision], (M, N, K, bn, ks)
            assert torch.equal(C1, C2), (M, N, K, bn, ks)
        C0, _ = sdvg_b200.gemm(A, W, b, relu=True, precision=precision)            # automatic plan (split-K when it pays)
        assert relerr(C0, want) < TOL[precision], (M, N, K)


@pytest.mark.parametrize("precision", ["fp16", "bf16", "fp32"])
@pytest.mark.parametrize("shape", [(40, 2048, 2048), (5, 2048, 2048), (48, 520, 1024), (17, 96, 4096), (1, 64, 192)])
def test_small_batch_gemm_compact_stages(precision, shape):
    """M <= 48 token rows: the one-CTA kernel's narrow tiles keep 48-row A stages (the MMA reads 128 rows from the stage
    base, the surplus rows feed accumulator rows nobody stores) in a ring of up to 16 stages, and issue their MMAs in
    batches.  Against float64, and bit-identical to the same rows computed inside a 128-row problem (full-height stages)."""
    import sdvg_b200
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M * 11 + N)
    A = torch.randn(128, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g)
    for bn in (32, 64):
        if bn > 32 and bn > N:
            continue
        small, _ = sdvg_b200.gemm(A[:M].contiguous(), W, b, precision=precision, block_n=bn)
        full, _ = sdvg_b200.gemm(A, W, b, precision=precision, block_n=bn)
        assert relerr(small, ref(A[:M], W, b, False, precision)) < TOL[precision]
        assert torch.equal(small, full[:M])
