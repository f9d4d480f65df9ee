"""GPU: parity of the CUDA path (through the Python mirror -> ctypes -> libsdvg C ABI) against the golden
outputs of the unmodified reference (tests/golden, made by oracle/make_golden.py) and against the oracle on
the same seeded inputs.

Tolerances (BASELINE.json north_star): per predicted frame max|ours-ref| / max|ref|
  fp32 modes  <= 1e-4 free-running          ("fp32" = split fp16x2 operands on tensor cores; "fp32_simt" = CUDA cores)
  16-bit mode <= 5e-3 teacher-forced        ("mixed" = fp16 operands, embedding + layer-0 QKV in split precision)
"""
import pytest
import torch

import sdvg_b200
from conftest import load_golden, ref_model_from_golden
from oracle import functional as F
from oracle import rollout as R

pytestmark = pytest.mark.gpu
TOL32, TOL16 = 1e-4, 5e-3
DEV = "cuda"


def maxrel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max())


def ours_from(g, precision):
    ref = ref_model_from_golden(g)
    d, H, Le, Ld, E = (int(v) for v in g["arch"])
    m = sdvg_b200.Transformer(0, d, H, Le, Ld, 0.1, frame_size={256: 64, 1024: 128}[E], precision=precision)
    m.load_state_dict(ref.state_dict())
    return m.eval().to(DEV), ref


@pytest.mark.parametrize("precision,tol", [("fp32_simt", TOL32), ("fp32", TOL32), ("mixed", TOL16), ("fp16", TOL16)])
def test_forward_golden_tiny(precision, tol):
    """forward(src, tgt, tgt_mask): S_src=6/S_tgt=5 (the trainer's shapes), mask tensor / None / causal, B=64."""
    g = load_golden("tiny_forward")
    m, ref = ours_from(g, precision)
    src, tgt, x64 = g["src"].to(DEV), g["tgt"].to(DEV), g["x64"].to(DEV)
    with torch.no_grad():
        assert maxrel(m(src, tgt, ref.get_tgt_mask(5)), g["out_causal"]) < tol      # CPU mask tensor, like predict.py:24
        assert maxrel(m(src, tgt, ref.get_tgt_mask(5).to(DEV)), g["out_causal"]) < tol
        assert maxrel(m(src, tgt, "causal"), g["out_causal"]) < tol
        assert maxrel(m(src, tgt), g["out_nomask"]) < tol
        assert maxrel(m(src, src, ref.get_tgt_mask(6)), g["out_same"]) < tol
        assert maxrel(m(src, src.clone(), ref.get_tgt_mask(6)), g["out_same"]) < tol  # src != tgt pointers: no dedupe
        out = m(x64, x64, "causal")
        assert tuple(out.shape) == (2, 64, 256)                                       # (S_tgt, B, E)
        assert maxrel(out, g["out_b64"]) < tol


def test_forward_shape_errors_and_b65():
    g = load_golden("tiny_forward")
    m, _ = ours_from(g, "fp32")
    x = torch.zeros(65, 2, 256, device=DEV)
    with pytest.raises(RuntimeError, match=r"size of tensor a \(65\)"):
        m(x, x)
    out = m(x, x, "causal", pe_index=torch.arange(65) % 64)                            # extension: explicit PE rows
    assert tuple(out.shape) == (2, 65, 256)
    with pytest.raises(RuntimeError):
        m(x[:2, :, :100], x[:2, :, :100])


def test_pe_indexed_by_batch_position():
    """the same clip at batch slot 0 vs 2 gives different outputs; pe_index reproduces it."""
    g = load_golden("tiny_forward")
    m, ref = ours_from(g, "fp32")
    src = g["src"].to(DEV)
    with torch.no_grad():
        full = m(src, src, "causal")
        alone = m(src[2:3], src[2:3], "causal", pe_index=torch.tensor([2]))
        wrong = m(src[2:3], src[2:3], "causal")
    assert maxrel(alone[:, 0], full[:, 2].cpu()) < 1e-5
    assert maxrel(wrong[:, 0], full[:, 2].cpu()) > 1e-3


@pytest.mark.parametrize("precision,tol", [("fp32_simt", TOL32), ("fp32", TOL32), ("mixed", TOL16)])
def test_forward_golden_d256(precision, tol):
    g = load_golden("d256_forward")
    m, _ = ours_from(g, precision)
    x = g["x"].to(DEV)
    with torch.no_grad():
        assert maxrel(m(x, x, "causal"), g["out"]) < tol


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32"])
def test_rollout_golden_small_free_running(precision):
    g = load_golden("small_rollout")
    m, _ = ours_from(g, precision)
    ctx = g["ctx"].to(DEV)
    assert R.max_rel_per_frame(sdvg_b200.rollout(m, ctx, 4, 5).cpu(), g["free5"]).max() < TOL32
    assert R.max_rel_per_frame(sdvg_b200.rollout(m, ctx, 3, 10).cpu(), g["free10"]).max() < TOL32
    fa = sdvg_b200.rollout(m, g["frames"].to(DEV), 4, 5, use_sos=True).cpu()          # literal predict.py sequence
    assert R.max_rel_per_frame(fa, g["faithful"]).max() < TOL32
    b1 = sdvg_b200.rollout(m, g["frames"][:1].to(DEV), 2, 5, use_sos=True).cpu()      # the reference's B=1 case
    assert R.max_rel_per_frame(b1, g["faithful_b1"]).max() < TOL32
    # predict(): clip 0, last position (prediction/predict.py:42)
    p = sdvg_b200.predict(m, ctx[:, -5:])
    assert tuple(p.shape) == (256,)
    assert maxrel(p, g["free5"][0, 0]) < TOL32


@pytest.mark.parametrize("precision", ["mixed", "fp16"])
def test_rollout_golden_small_teacher_forced_16bit(precision):
    g = load_golden("small_rollout")
    m, _ = ours_from(g, precision)
    out = sdvg_b200.rollout(m, g["ctx"].to(DEV), 4, 5, teacher=g["free5"].to(DEV)).cpu()
    assert R.max_rel_per_frame(out, g["free5"]).max() < TOL16


def test_rollout_scales_and_host_path():
    g = load_golden("small_rollout")
    m, _ = ours_from(g, "fp32")
    ctx = g["ctx"].to(DEV)
    base = sdvg_b200.rollout(m, ctx, 2, 5)
    s = sdvg_b200.LATENT_SCALE
    scaled = sdvg_b200.rollout(m, ctx / s, 2, 5, scale_in=s, scale_out=1.0 / s)       # utils/sd_utils.py:143,159
    assert maxrel(scaled * s, base.cpu()) < 1e-5
    host = sdvg_b200.rollout_from_host(m, g["ctx"], 2, 5)
    assert not host.is_cuda and torch.equal(host, base.cpu())


def test_rollout_is_deterministic_and_batch_invariant():
    g = load_golden("small_rollout")
    m, _ = ours_from(g, "fp32")
    ctx = torch.randn(200, 6, 256, generator=torch.Generator().manual_seed(3)).to(DEV)
    a = sdvg_b200.rollout(m, ctx, 2, 5)
    b = sdvg_b200.rollout(m, ctx, 2, 5)
    assert torch.equal(a, b)
    # a clip's result depends only on its own data and its PE row: shards == full batch, bit for bit
    parts = [sdvg_b200.rollout(m, ctx[s:e], 2, 5, pe_index=sdvg_b200.pe_index_for(s, e)) for s, e in ((0, 64), (64, 128), (128, 200))]
    assert torch.equal(torch.cat(parts), a)


def test_b_gt_64_equals_reference_in_chunks():
    """B=130 in one call == the reference run in chunks of <= 64 clips (clip i sees PE[i mod 64])."""
    g = load_golden("small_rollout")
    m, ref = ours_from(g, "fp32")
    ctx = torch.randn(130, 6, 256, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = R.chunked(lambda c: R.rollout_ref(ref, c, 2, 5), ctx)
    got = sdvg_b200.rollout(m, ctx.to(DEV), 2, 5).cpu()
    assert R.max_rel_per_frame(got, want).max() < TOL32


@pytest.mark.parametrize("name", ["c1_rollout", "c4_rollout"])
def test_full_size_configs_golden(name):
    """BASELINE configs at full width: C1/C2 arch (d2048 4e/8d E256) and C4 (d2048 6e/6d E1024)."""
    g = load_golden(name)
    n = g["free5"].shape[1]
    m, _ = ours_from(g, "fp32")
    ctx = g["ctx"].to(DEV)
    assert R.max_rel_per_frame(sdvg_b200.rollout(m, ctx, n, 5).cpu(), g["free5"]).max() < TOL32
    if "free10" in g:
        assert R.max_rel_per_frame(sdvg_b200.rollout(m, ctx, 1, 10).cpu(), g["free10"]).max() < TOL32
    m.set_precision("mixed")
    tf = sdvg_b200.rollout(m, ctx, n, 5, teacher=g["free5"].to(DEV)).cpu()
    assert R.max_rel_per_frame(tf, g["free5"]).max() < TOL16


def test_bench_size_batch_properties():
    """B=1024 clips (BASELINE configs[1] size) on the C1/C2 architecture: no oracle at this size, so check
    size-independent properties - the first 8 clips equal the golden 8-clip run, and a 64-aligned shard equals
    the same clips inside the big batch."""
    g = load_golden("c1_rollout")
    m, _ = ours_from(g, "fp32")
    gen = torch.Generator().manual_seed(99)
    ctx = torch.randn(1024, 10, 256, generator=gen)
    ctx[:8] = g["ctx"]
    ctx = ctx.to(DEV)
    out = sdvg_b200.rollout(m, ctx, 2, 5)
    assert torch.isfinite(out).all()
    assert R.max_rel_per_frame(out[:8].cpu(), g["free5"][:, :2]).max() < TOL32
    shard = sdvg_b200.rollout(m, ctx[512:640], 2, 5, pe_index=sdvg_b200.pe_index_for(512, 640))
    assert R.max_rel_per_frame(shard.cpu(), out[512:640].cpu()).max() < 1e-5


def test_token_local_cache_is_exact(monkeypatch):
    """The rollout computes the embedding and the layer-0 Q/K/V of each frame once (token-local caches, the only
    exact K/V reuse this model allows - SURVEY.md fact 5).  Same results as recomputing the whole window."""
    g = load_golden("small_rollout")
    ctx = torch.randn(70, 7, 256, generator=torch.Generator().manual_seed(8)).to(DEV)
    teacher = torch.randn(70, 5, 256, generator=torch.Generator().manual_seed(9)).to(DEV)
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("SDVG_CACHE", flag)          # read when the handle is created
        for prec in ("fp32", "mixed"):
            m, _ = ours_from(g, prec)
            outs[flag, prec, "free"] = sdvg_b200.rollout(m, ctx, 5, 5)
            outs[flag, prec, "w3"] = sdvg_b200.rollout(m, ctx, 5, 3)
            outs[flag, prec, "grow"] = sdvg_b200.rollout(m, ctx[:, :2], 5, 6)      # window grows from 2 to 6 tokens
            outs[flag, prec, "tf"] = sdvg_b200.rollout(m, ctx, 5, 5, teacher=teacher)
    for (flag, prec, kind), v in outs.items():
        if flag == "1":
            ref = outs["0", prec, kind]
            assert R.max_rel_per_frame(v.cpu(), ref.cpu()).max() < 1e-6, (prec, kind)


def test_sibling_variants():
    """predict_diff residual prediction (prediction/predict_diff.py:33), the one-shot unmasked
    call of predict_future.py:156, TransformerFuture checkpoints, the Identity baseline."""
    g = load_golden("small_rollout")
    m, ref = ours_from(g, "fp32")
    ctx = g["ctx"].to(DEV)
    with torch.no_grad():
        want = R.rollout_ref(ref, g["ctx"], 3, 5, residual=True)
        got = sdvg_b200.rollout(m, ctx, 3, 5, residual=True).cpu()
        assert R.max_rel_per_frame(got, want).max() < TOL32
        wf = R.rollout_faithful(ref, g["frames"], 3, residual=True)
        gf = sdvg_b200.rollout(m, g["frames"].to(DEV), 3, 5, use_sos=True, residual=True).cpu()
        assert R.max_rel_per_frame(gf, wf).max() < TOL32
        x = g["ctx"][:, -5:]
        assert maxrel(sdvg_b200.predict_diff(m, x.to(DEV)), R.predict_diff_ref(ref, x)[0]) < TOL32
        fut = ref(x, x, None).permute(1, 0, 2)                                       # predict_future.py:156
        assert maxrel(m(x.to(DEV), x.to(DEV), None).permute(1, 0, 2), fut) < TOL32
        assert maxrel(sdvg_b200.predict_future(m, x.to(DEV)), fut[0, -1]) < TOL32
    d, H, Le, Ld, E = (int(v) for v in g["arch"])
    mf = sdvg_b200.TransformerFuture(0, d, H, Le, Ld, 0.1, frame_size=64, frames_to_predict=5, precision="fp32")
    sd = dict(ref.state_dict()); sd["learned_tgt"] = torch.zeros(1, 5, E)            # transformer_future.py:46-47
    mf.load_state_dict(sd)
    mf = mf.eval().to(DEV)
    with torch.no_grad():
        assert maxrel(mf(x.to(DEV), x.to(DEV), None).permute(1, 0, 2), fut) < TOL32
    ident = sdvg_b200.Identity()
    assert torch.equal(ident(ctx, ctx), ctx[:, -1:])


def test_cuda_graph_replay_matches_eager(monkeypatch):
    """The third call with identical arguments replays a captured CUDA graph (second call captures): same bits,
    same launch accounting; new input values in the same buffer are honoured; SDVG_GRAPH=0 stays eager."""
    g = load_golden("small_rollout")
    m, _ = ours_from(g, "fp32")
    ctx = g["ctx"].to(DEV).clone()
    out = torch.empty(4, 4, 256, device=DEV)
    n0 = m.launch_count()
    a = sdvg_b200.rollout(m, ctx, 4, 5, out=out).clone(); n1 = m.launch_count()
    b = sdvg_b200.rollout(m, ctx, 4, 5, out=out).clone(); n2 = m.launch_count()     # capture + first replay
    c = sdvg_b200.rollout(m, ctx, 4, 5, out=out).clone(); n3 = m.launch_count()     # replay
    assert torch.equal(a, b) and torch.equal(a, c)
    assert n2 - n1 == n3 - n2 > 0 and n1 - n0 >= n2 - n1      # the first call also packs the weight planes
    assert R.max_rel_per_frame(c.cpu(), g["free5"]).max() < TOL32
    ctx.mul_(0.5)                                                                      # same pointer, new contents
    d = sdvg_b200.rollout(m, ctx, 4, 5, out=out).clone()
    monkeypatch.setenv("SDVG_GRAPH", "0")
    m2, _ = ours_from(g, "fp32")
    e1 = sdvg_b200.rollout(m2, ctx, 4, 5)
    e2 = sdvg_b200.rollout(m2, ctx, 4, 5)
    e3 = sdvg_b200.rollout(m2, ctx, 4, 5)
    assert torch.equal(d, e3) and torch.equal(e1, e3) and torch.equal(e2, e3)


@pytest.mark.parametrize("name", ["11_27_ucf_final", "11_19_wallpushups_all_losses_test"])
def test_other_baseline_architectures_vs_oracle(name):
    """BASELINE configs C3 (d2048 H8 4e/8d, E=1024) and C5's architecture (d1024 H16 12e/12d, E=1024, head dim 64):
    no golden file - the oracle (reference arithmetic on CPU, same seeded weights) is run here on a small batch."""
    from oracle.ref_module import RefTransformer
    c = sdvg_b200.CONFIGS[name]
    torch.manual_seed(0)
    ref = RefTransformer(0, c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"], 0.1,
                         frame_size=c["frame_size"]).eval()
    m = sdvg_b200.Transformer(0, c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"], 0.1,
                              frame_size=c["frame_size"], precision="fp32")
    m.load_state_dict(ref.state_dict())
    m = m.eval().to(DEV)
    ctx = torch.randn(4, 6, 1024, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        want = R.rollout_ref(ref, ctx, 2, 5)
        fwd = ref(ctx[:, :6], ctx[:, :5], ref.get_tgt_mask(5))                  # trainer shapes: S_src=6, S_tgt=5
    assert R.max_rel_per_frame(sdvg_b200.rollout(m, ctx.to(DEV), 2, 5).cpu(), want).max() < TOL32
    assert maxrel(m(ctx[:, :6].to(DEV), ctx[:, :5].contiguous().to(DEV), "causal"), fwd) < TOL32
    m.set_precision("mixed")
    tf = sdvg_b200.rollout(m, ctx.to(DEV), 2, 5, teacher=want.to(DEV)).cpu()
    assert R.max_rel_per_frame(tf, want).max() < TOL16


def test_tensor_core_attention_windows(monkeypatch):
    """Head dim 256 (d512, H2) in the 16-bit mode routes attention through the mma.sync kernels (<= 8 and <= 16
    tokens).  Windows 5, 6 (SOS), 10 and 16 against the oracle (teacher-forced, 5e-3) and against the warp-level
    kernels (SDVG_ATTN_MMA=0) - the two differ only in rounding P to 16 bits."""
    from oracle.ref_module import RefTransformer
    torch.manual_seed(5)
    ref = RefTransformer(0, 512, 2, 1, 2, 0.1, frame_size=64).eval()
    ctx = torch.randn(9, 16, 256, generator=torch.Generator().manual_seed(22))
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SDVG_ATTN_MMA", flag)
        m = sdvg_b200.Transformer(0, 512, 2, 1, 2, 0.1, frame_size=64, precision="mixed")
        m.load_state_dict(ref.state_dict())
        m = m.eval().to(DEV)
        for C, W in ((5, 5), (10, 10), (16, 16), (12, 7)):
            with torch.no_grad():
                want = R.rollout_ref(ref, ctx[:, :C], 3, W)
            got = sdvg_b200.rollout(m, ctx[:, :C].contiguous().to(DEV), 3, W, teacher=want.to(DEV)).cpu()
            assert R.max_rel_per_frame(got, want).max() < TOL16, (flag, C, W)
            outs[flag, C, W] = got
        with torch.no_grad():
            want = R.rollout_faithful(ref, ctx[:, :5], 3)
        got = sdvg_b200.rollout(m, ctx[:, :5].contiguous().to(DEV), 3, use_sos=True, teacher=want.to(DEV)).cpu()
        assert R.max_rel_per_frame(got, want).max() < TOL16, flag
        outs[flag, "sos"] = got
        with torch.no_grad():
            fwd = ref(ctx[:, :11], ctx[:, :10], ref.get_tgt_mask(10))
        o = m(ctx[:, :11].contiguous().to(DEV), ctx[:, :10].contiguous().to(DEV), "causal")
        assert maxrel(o, fwd) < TOL16, flag
        outs[flag, "fwd"] = o.cpu()
    for k in [k[1:] for k in outs if k[0] == "1"]:
        assert maxrel(outs[("1",) + k], outs[("0",) + k]) < 2e-3, k


def test_stacked_cross_attention_projection_is_bit_identical(monkeypatch):
    """16-bit modes project the cross-attention K|V of all decoder layers in one GEMM against stacked weights (they all
    read the same encoder memory).  Same bits as one GEMM per layer (SDVG_BATCH_CROSS=0), for pruned and full passes."""
    g = load_golden("small_rollout")
    ctx = torch.randn(70, 7, 256, generator=torch.Generator().manual_seed(23)).to(DEV)
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SDVG_BATCH_CROSS", flag)
        for prec in ("mixed", "fp16"):
            m, _ = ours_from(g, prec)
            outs[flag, prec, "roll"] = sdvg_b200.rollout(m, ctx, 3, 5)
            outs[flag, prec, "fwd"] = m(ctx[:64, :6].contiguous(), ctx[:64, :5].contiguous(), "causal")
    for k in [k[1:] for k in outs if k[0] == "1"]:
        assert torch.equal(outs[("1",) + k], outs[("0",) + k]), k


@pytest.mark.parametrize("d,H", [(96, 4), (32, 4), (192, 2)])
def test_odd_model_widths(d, H):
    """Widths that are not multiples of 64 (operand planes padded to 64 columns, partial K blocks zero-filled by TMA)
    and head sizes 24 / 8 / 96 (generic attention kernels): forward and rollout against the oracle, fp32 and mixed."""
    from oracle.ref_module import RefTransformer
    torch.manual_seed(31)
    ref = RefTransformer(0, d, H, 2, 2, 0.1, frame_size=64).eval()
    ctx = torch.randn(5, 7, 256, generator=torch.Generator().manual_seed(32))
    with torch.no_grad():
        want = R.rollout_ref(ref, ctx, 3, 5)
        fwd = ref(ctx[:, :6], ctx[:, :5], ref.get_tgt_mask(5))
    m = sdvg_b200.Transformer(0, d, H, 2, 2, 0.1, frame_size=64, precision="fp32")
    m.load_state_dict(ref.state_dict())
    m = m.eval().to(DEV)
    assert R.max_rel_per_frame(sdvg_b200.rollout(m, ctx.to(DEV), 3, 5).cpu(), want).max() < TOL32
    assert maxrel(m(ctx[:, :6].contiguous().to(DEV), ctx[:, :5].contiguous().to(DEV), "causal"), fwd) < TOL32
    m.set_precision("mixed")
    tf = sdvg_b200.rollout(m, ctx.to(DEV), 3, 5, teacher=want.to(DEV)).cpu()
    assert R.max_rel_per_frame(tf, want).max() < TOL16


def test_edge_shapes():
    """Single clip, single token, window longer than the history, maximum batch of the reference (64), 32-token window."""
    g = load_golden("small_rollout")
    m, ref = ours_from(g, "fp32")
    with torch.no_grad():
        x1 = torch.randn(1, 1, 256, generator=torch.Generator().manual_seed(2))
        assert maxrel(m(x1.to(DEV), x1.to(DEV), "causal"), ref(x1, x1, ref.get_tgt_mask(1))) < TOL32
        c1 = torch.randn(1, 1, 256, generator=torch.Generator().manual_seed(3))           # C=1, window 5: growing window
        assert R.max_rel_per_frame(sdvg_b200.rollout(m, c1.to(DEV), 3, 5).cpu(), R.rollout_ref(ref, c1, 3, 5)).max() < TOL32
        x32 = torch.randn(2, 32, 256, generator=torch.Generator().manual_seed(4))          # S = 32 (kernel limit)
        assert maxrel(m(x32.to(DEV), x32.to(DEV), "causal"), ref(x32, x32, ref.get_tgt_mask(32))) < TOL32
        x64 = torch.randn(64, 3, 256, generator=torch.Generator().manual_seed(5))
        assert maxrel(m(x64.to(DEV), x64.to(DEV), None), ref(x64, x64, None)) < TOL32
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 33, 256, device=DEV), torch.zeros(2, 33, 256, device=DEV))       # beyond max tokens
    with pytest.raises(RuntimeError):
        sdvg_b200.rollout(m, torch.zeros(0, 5, 256, device=DEV), 2, 5)                     # empty batch


def test_last_layer_pruning_is_exact(monkeypatch):
    """Rollouts compute the last decoder layer for the last token of each clip only (predict.py:42 keeps nothing
    else).  Same results as the full layer, for fixed, growing and odd windows and for 16-bit modes."""
    g = load_golden("small_rollout")
    ctx = torch.randn(70, 7, 256, generator=torch.Generator().manual_seed(18)).to(DEV)
    outs = {}
    # (the layer that feeds the pruned one keeps a LayerNorm kernel where the unpruned model folds it into the GEMMs: with
    # folding the two runs differ by 16-bit rounding, not by the pruning - keep it out of this comparison)
    monkeypatch.setenv("SDVG_LN_FOLD", "0")
    for flag in ("0", "1"):
        monkeypatch.setenv("SDVG_PRUNE", flag)
        for prec in ("fp32", "mixed"):
            m, _ = ours_from(g, prec)
            outs[flag, prec, "w5"] = sdvg_b200.rollout(m, ctx, 4, 5)
            outs[flag, prec, "w7"] = sdvg_b200.rollout(m, ctx, 3, 7)
            outs[flag, prec, "grow"] = sdvg_b200.rollout(m, ctx[:, :2], 5, 6)
            outs[flag, prec, "sos"] = sdvg_b200.rollout(m, ctx[:, :5], 3, 5, use_sos=True)
    for (flag, prec, kind), v in outs.items():
        if flag == "1":
            assert R.max_rel_per_frame(v.cpu(), outs["0", prec, kind].cpu()).max() < 1e-6, (prec, kind)


def test_losses_match_reference_trainer_gpu():
    """Trainer.criterion / gradient_difference_loss / BiPatchNCE forward values (trainers/trainer.py:65-109,
    models/contrastive_loss.py) through sdvg_criterion vs the golden values of the unmodified reference."""
    g = load_golden("losses")
    close = lambda a, b: abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b)))
    for tag in ("f64", "f128", "p1"):
        x, y = g[f"{tag}.x"].to(DEV), g[f"{tag}.y"].to(DEV)
        P, B, F = (int(v) for v in g[f"{tag}.shape"])
        t = sdvg_b200.loss_terms(x, y, alpha=2, temperature=0.07)
        assert close(t["mse"], g[f"{tag}.mse"]) and close(t["l1"], g[f"{tag}.l1"])
        assert close(t["gdl"], g[f"{tag}.gdl2"]) and close(t["contrastive"], g[f"{tag}.nce"])
        assert close(sdvg_b200.gradient_difference_loss(x, y, 1), g[f"{tag}.gdl1"])
        nce = sdvg_b200.BiPatchNCE(N=B, T=P, h=F // 8, w=F // 8, temperature=0.07)
        assert close(nce(x.permute(1, 0, 2).reshape(-1, P, 4, F // 8, F // 8), y.permute(1, 0, 2).reshape(-1, P, 4, F // 8, F // 8)),
                     g[f"{tag}.nce"])
        assert close(sdvg_b200.criterion(False, True, False, use_contrastive=False)(x, y), g[f"{tag}.crit_l1"])
        assert close(sdvg_b200.criterion(True, False, True, 1, 2, True, 0.07, 0.1)(x, y), g[f"{tag}.crit_c5"])
        assert close(sdvg_b200.criterion(False, True, True, 0.5, 1, True, 0.1, 0.0258)(x, y), g[f"{tag}.crit_all"])
    assert sdvg_b200.criterion(use_mse=True, use_L1=True) is None


def test_validation_step_matches_reference_loop():
    """One validation_loop iteration (trainers/trainer.py:203-224): S_src = T+1 (with SOS), S_tgt = T, causal mask,
    loss on the last frames - vs the reference model + the oracle's loss restatement."""
    from oracle import losses as L
    g = load_golden("small_rollout")
    m, ref = ours_from(g, "fp32")
    frames = torch.randn(4, 5, 256, generator=torch.Generator().manual_seed(31)) * sdvg_b200.LATENT_SCALE
    batch = torch.cat([torch.full((4, 1, 256), sdvg_b200.SOS_VALUE), frames], 1)           # encode_batch(use_sos=True)
    with torch.no_grad():
        pred_ref = ref(batch, batch[:, :-1], ref.get_tgt_mask(5))
        want = L.criterion(True, False, True, 1, 2, True, 0.07, 0.1)(pred_ref[-5:], batch[:, 1:].permute(1, 0, 2)[-5:])
    loss_fn = sdvg_b200.criterion(True, False, True, 1, 2, True, 0.07, 0.1)
    loss, pred = sdvg_b200.validation_step(m, batch.to(DEV), 5, loss_fn)
    assert maxrel(pred, pred_ref) < TOL32
    assert abs(float(loss) - float(want)) <= 1e-4 * abs(float(want))


def test_bench_scale_mixed_teacher_forced_vs_chunked_oracle():
    """BASELINE configs[1] at bench scale in the bench's own precision: B = 1024 clips on the C1/C2 architecture,
    `mixed`, teacher-forced (the 5e-3 bar of north_star is defined teacher-forced), checked against the chunked oracle on
    a 128-clip subset spanning two PE chunks - clips 0..63 and 960..1023 - for 2 predicted frames."""
    g = load_golden("c1_rollout")
    m, ref = ours_from(g, "mixed")
    ctx = torch.randn(1024, 10, 256, generator=torch.Generator().manual_seed(99))
    sub = torch.cat([torch.arange(0, 64), torch.arange(960, 1024)])
    with torch.no_grad():
        want = R.chunked(lambda c: R.rollout_ref(ref, c, 2, 5), ctx[sub])          # two chunks of 64: PE rows 0..63 each
    teacher = torch.zeros(1024, 2, 256)
    teacher[sub] = want
    out = sdvg_b200.rollout(m, ctx.to(DEV), 2, 5, teacher=teacher.to(DEV)).cpu()
    assert torch.isfinite(out).all()
    assert R.max_rel_per_frame(out[sub], want).max() < TOL16


def test_c4_full_rollout_mixed():
    """The "wide" config (11_20_wallpushups_dim_2048, d2048 6e/6d E1024): every frame of the golden rollout in `mixed`,
    teacher-forced, and free-running in fp32."""
    g = load_golden("c4_rollout")
    n = g["free5"].shape[1]
    m, _ = ours_from(g, "mixed")
    ctx = g["ctx"].to(DEV)
    tf = sdvg_b200.rollout(m, ctx, n, 5, teacher=g["free5"].to(DEV)).cpu()
    errs = R.max_rel_per_frame(tf, g["free5"])
    assert errs.max() < TOL16, errs.tolist()
    # a longer teacher-forced run than the fixture holds: 10 frames against the oracle port on 4 clips
    ref = ref_model_from_golden(g)
    with torch.no_grad():
        want = R.rollout_ref(ref, g["ctx"][:4], 10, 5)
    out = sdvg_b200.rollout(m, ctx[:4], 10, 5, teacher=want.to(DEV)).cpu()
    errs = R.max_rel_per_frame(out, want)
    assert errs.max() < TOL16, errs.tolist()


def test_rollout_from_the_npy_latent_cache(tmp_path):
    """(f4) the reference's offline latent cache (utils/preprocess.py:27-33) feeds the hot path through host buffers."""
    g = load_golden("small_rollout")
    m, _ = ours_from(g, "fp32")
    dirs = []
    for b in range(g["ctx"].shape[0]):
        d = str(tmp_path / f"clip{b:02d}")
        sdvg_b200.save_latent_frames(d, g["ctx"][b])
        dirs.append(d)
    ctx = sdvg_b200.load_latent_clips(dirs)
    assert torch.equal(ctx, g["ctx"])
    out = sdvg_b200.rollout_from_host(m, ctx, 4, 5)
    assert R.max_rel_per_frame(out, g["free5"]).max() < TOL32


def test_layernorm_folded_into_the_gemms(monkeypatch):
    """Large batches in the 16-bit modes run without LayerNorm kernels between the sub-layers: the producing GEMM's
    epilogue leaves per-row partial sums, the consuming GEMM reads the pre-norm sums through gamma-scaled weights and
    applies rstd * (acc - mean * c) + b' (Engine::fold_*, SDVG_LN_FOLD).  Forced on at any batch size here: the golden
    forward / rollout values of the reference within the 16-bit tolerance, agreement with the unfolded path at rounding
    level, and no effect on the fp32 mode (never folded)."""
    g = load_golden("small_rollout")
    ctx = g["ctx"].to(DEV)
    n = g["free5"].shape[1]
    outs = {}
    for fold in ("0", "1"):
        monkeypatch.setenv("SDVG_LN_FOLD", fold)
        monkeypatch.setenv("SDVG_LN_FOLD_MIN", "1")
        for prec in ("mixed", "fp16", "bf16", "fp32"):
            m, _ = ours_from(g, prec)
            tf = sdvg_b200.rollout(m, ctx, n, 5, teacher=g["free5"].to(DEV)).cpu()
            outs[fold, prec] = tf
            tol = {"fp32": TOL32, "bf16": 4e-2}.get(prec, TOL16)
            assert R.max_rel_per_frame(tf, g["free5"]).max() < tol, (fold, prec)
            big = torch.randn(130, 6, 256, generator=torch.Generator().manual_seed(7)).to(DEV)       # several row tiles, B > 64
            pe = (torch.arange(130) % 64).to(torch.int32).to(DEV)
            outs[fold, prec, "big"] = sdvg_b200.rollout(m, big, 2, 5, pe_index=pe).cpu()
    assert torch.equal(outs["0", "fp32"], outs["1", "fp32"])
    for prec in ("mixed", "fp16"):
        assert not torch.equal(outs["0", prec], outs["1", prec])                         # the folded path really ran
        assert R.max_rel_per_frame(outs["1", prec, "big"], outs["0", prec, "big"])[0] < TOL16


def test_folded_layernorm_statistics_inline_vs_statistics_kernel(monkeypatch):
    """Up to 128 token rows the consumers of a folded LayerNorm add the producer's partial sums themselves
    (Epilogue::stat_in, SDVG_STATS_INLINE_MAX); above that a statistics kernel does.  Same partial sums, different summation
    order: the two paths agree to fp32 rounding through the 16-bit pipeline, both within the golden tolerance, and the
    small-batch result does not depend on which rollouts ran before (double-buffered slots)."""
    g = load_golden("small_rollout")
    ctx = g["ctx"].to(DEV)
    n = g["free5"].shape[1]
    outs = {}
    for inline in ("128", "0"):
        monkeypatch.setenv("SDVG_STATS_INLINE_MAX", inline)
        m, _ = ours_from(g, "mixed")
        a = sdvg_b200.rollout(m, ctx, n, 5, teacher=g["free5"].to(DEV)).cpu()
        b = sdvg_b200.rollout(m, ctx, n, 5, teacher=g["free5"].to(DEV)).cpu()
        assert torch.equal(a, b), inline
        assert R.max_rel_per_frame(a, g["free5"]).max() < TOL16, inline
        outs[inline] = a
    assert R.max_rel_per_frame(outs["128"], outs["0"]).max() < 2e-3
