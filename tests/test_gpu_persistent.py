"""GPU: the persistent small-batch kernel (csrc/persistent.cuh, SDVG_PK=1) - one launch per forward pass / rollout when
clips x tokens <= 128 - against the golden outputs of the unmodified reference and against the per-kernel path.

Tolerances as in test_gpu_parity.py: fp32 mode <= 1e-4 free-running, 16-bit modes <= 5e-3 teacher-forced."""
import pytest
import torch

import sdvg_b200
from conftest import load_golden, ref_model_from_golden
from oracle import rollout as R

pytestmark = pytest.mark.gpu
TOL32, TOL16 = 1e-4, 5e-3
DEV = "cuda"


def maxrel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max())


def model_from(g, precision, monkeypatch, pk=True):
    monkeypatch.setenv("SDVG_PK", "1" if pk else "0")      # read when the engine is created
    ref = ref_model_from_golden(g)
    d, H, Le, Ld, E = (int(v) for v in g["arch"])
    m = sdvg_b200.Transformer(0, d, H, Le, Ld, 0.1, frame_size={256: 64, 1024: 128}[E], precision=precision)
    m.load_state_dict(ref.state_dict())
    m = m.eval().to(DEV)
    m.engine(torch.device(DEV, 0))
    return m, ref


def ref64(A, W, bias, relu, precision):
    if precision in ("fp16", "bf16"):
        dt = torch.float16 if precision == "fp16" else torch.bfloat16
        A, W = A.to(dt), W.to(dt)
    y = A.double() @ W.double().t() + bias.double()
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_persistent_gemm_matches_float64(precision):
    """One-op programs: ragged M / N / K, tiles of every height class (T <= 32, <= 64, > 64), K blocks that leave CTAs
    of a cluster without work, two token groups (M > 64); deterministic from launch to launch."""
    for (M, N, K) in ((40, 2048, 2048), (8, 256, 2048), (48, 2048, 256), (80, 6144, 2048), (128, 512, 1024), (5, 200, 512),
                      (77, 96, 32), (1, 32, 8), (40, 4200, 1024)):
        g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
        A = torch.randn(M, K, device=DEV, generator=g)
        W = torch.randn(N, K, device=DEV, generator=g) * 0.05
        b = torch.randn(N, device=DEV, generator=g)
        C1, _ = sdvg_b200.gemm(A, W, b, relu=True, precision=precision, block_n=9999)
        C2, _ = sdvg_b200.gemm(A, W, b, relu=True, precision=precision, block_n=9999, iters=3)
        assert maxrel(C1, ref64(A, W, b, True, precision)) < 1e-5, (M, N, K)
        assert torch.equal(C1, C2), (M, N, K)


@pytest.mark.parametrize("precision,tol", [("fp32", TOL32), ("mixed", TOL16), ("fp16", TOL16)])
def test_persistent_forward_golden(precision, tol, monkeypatch):
    g = load_golden("tiny_forward")
    m, ref = model_from(g, precision, monkeypatch)
    src, tgt = g["src"].to(DEV), g["tgt"].to(DEV)
    with torch.no_grad():
        n0 = m.launch_count()
        out = m(src, tgt, ref.get_tgt_mask(5))
        assert m.launch_count() - n0 == 1                       # the whole pass is one launch
        assert maxrel(out, g["out_causal"]) < tol
        assert maxrel(m(src, tgt, ref.get_tgt_mask(5).to(DEV)), g["out_causal"]) < tol      # additive mask tensor
        assert maxrel(m(src, tgt), g["out_nomask"]) < tol
        assert maxrel(m(src, src, "causal"), g["out_same"]) < tol
    g2 = load_golden("d256_forward")
    m2, _ = model_from(g2, precision, monkeypatch)
    x = g2["x"].to(DEV)
    with torch.no_grad():
        assert maxrel(m2(x, x, "causal"), g2["out"]) < tol


def test_persistent_rollout_small_golden(monkeypatch):
    """Sliding window (token-local caches, last-layer pruning), window 10, the literal predict.py sequence, B = 1."""
    g = load_golden("small_rollout")
    m, _ = model_from(g, "fp32", monkeypatch)
    ctx = g["ctx"].to(DEV)
    n0 = m.launch_count()
    out = sdvg_b200.rollout(m, ctx, 4, 5)
    assert m.launch_count() - n0 == 1
    assert R.max_rel_per_frame(out.cpu(), g["free5"]).max() < TOL32
    assert R.max_rel_per_frame(sdvg_b200.rollout(m, ctx, 3, 10).cpu(), g["free10"]).max() < TOL32
    fa = sdvg_b200.rollout(m, g["frames"].to(DEV), 4, 5, use_sos=True).cpu()
    assert R.max_rel_per_frame(fa, g["faithful"]).max() < TOL32
    b1 = sdvg_b200.rollout(m, g["frames"][:1].to(DEV), 2, 5, use_sos=True).cpu()      # the reference's own regime
    assert R.max_rel_per_frame(b1, g["faithful_b1"]).max() < TOL32
    again = sdvg_b200.rollout(m, ctx, 4, 5)                                            # cached program, same bits
    assert torch.equal(again, out)
    m16, _ = model_from(g, "mixed", monkeypatch)
    tf = sdvg_b200.rollout(m16, ctx, 4, 5, teacher=g["free5"].to(DEV)).cpu()
    assert R.max_rel_per_frame(tf, g["free5"]).max() < TOL16


def test_persistent_c1_rollout_golden(monkeypatch):
    """BASELINE configs[0] at full width (d2048 4e/8d E256, B=8) through the persistent kernel."""
    g = load_golden("c1_rollout")
    n = g["free5"].shape[1]
    m, _ = model_from(g, "fp32", monkeypatch)
    ctx = g["ctx"].to(DEV)
    n0 = m.launch_count()
    out = sdvg_b200.rollout(m, ctx, n, 5)
    assert m.launch_count() - n0 == 1
    assert R.max_rel_per_frame(out.cpu(), g["free5"]).max() < TOL32
    if "free10" in g:
        assert R.max_rel_per_frame(sdvg_b200.rollout(m, ctx, 1, 10).cpu(), g["free10"]).max() < TOL32
    m.set_precision("mixed")
    m.engine(torch.device(DEV, 0))
    tf = sdvg_b200.rollout(m, ctx, n, 5, teacher=g["free5"].to(DEV)).cpu()
    assert R.max_rel_per_frame(tf, g["free5"]).max() < TOL16


def test_persistent_equals_per_kernel_path(monkeypatch):
    """Both paths run the same sequencing code (Engine::run_model); they differ in summation order only."""
    g = load_golden("small_rollout")
    ctx = torch.randn(9, 7, 256, generator=torch.Generator().manual_seed(8)).to(DEV)
    for prec, tol in (("fp32", 2e-5), ("mixed", 5e-3)):
        a, _ = model_from(g, prec, monkeypatch, pk=True)
        b, _ = model_from(g, prec, monkeypatch, pk=False)
        teacher = torch.randn(9, 3, 256, generator=torch.Generator().manual_seed(9)).to(DEV)
        for kw in (dict(), dict(residual=True), dict(teacher=teacher), dict(pe_index=0)):
            x = sdvg_b200.rollout(a, ctx, 3, 5, **kw)
            y = sdvg_b200.rollout(b, ctx, 3, 5, **kw)
            assert R.max_rel_per_frame(x.cpu(), y.cpu()).max() < tol, (prec, kw)
    # a batch too large for the persistent kernel silently takes the per-kernel path of the same handle
    big = torch.randn(40, 7, 256, generator=torch.Generator().manual_seed(10)).to(DEV)
    n0 = a.launch_count()
    sdvg_b200.rollout(a, big, 2, 5)
    assert a.launch_count() - n0 > 1
