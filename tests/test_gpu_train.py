"""GPU: the training step (trainers/trainer.py:123-162 - teacher-forced forward, criterion, backward, Adam) through
libsdvg's C ABI against (a) golden values of the unmodified reference under torch autograd
(tests/golden/train_step.npz, oracle/make_golden_train.py) and (b) the oracle on the same seeded inputs.

Tolerance: fp32 mode (split-precision tensor-core GEMMs), per parameter tensor max|ours - ref| <= 1e-4 * max|ref|
for gradients; losses to 1e-5 relative; weights after two Adam steps to 2e-3 of the total parameter change."""
import numpy as np
import pytest
import torch

import sdvg_b200
from conftest import load_golden, sd_checksum
from oracle import train as OT
from oracle.ref_module import RefTransformer

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOLG = 1e-4


def sample(t, n=64):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long().to(f.device)
    return f[idx]


def build_pair(d, H, Le, Ld, seed, frame_size=64, precision="fp32"):
    torch.manual_seed(seed)
    ref = RefTransformer(0, d, H, Le, Ld, 0.0, frame_size=frame_size)
    m = sdvg_b200.Transformer(0, d, H, Le, Ld, 0.0, frame_size=frame_size, precision=precision)
    m.load_state_dict(ref.state_dict())
    return m.to(DEV), ref


def relu_margin(ref, batch):
    """Smallest |pre-activation| / rms over every FFN hidden unit of a teacher-forced forward.  A ReLU whose input is
    within rounding distance of zero may open in one fp32 implementation and close in another; its gradient row then
    differs by O(1) - a discontinuity of the model, not an error of either side - so the large-batch test picks an
    input where no unit sits on the kink."""
    zs = []
    hooks = [mod.register_forward_hook(lambda _m, _i, out: zs.append(out.detach()))
             for name, mod in ref.named_modules() if name.endswith("linear1")]
    with torch.no_grad():
        ref.eval()
        ref(batch, batch[:, :-1], ref.get_tgt_mask(batch.size(1) - 1))
        ref.train()
    for h in hooks:
        h.remove()
    return min(float(z.abs().min() / z.pow(2).mean().sqrt()) for z in zs)


CASES = {
    "c5": dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07,
               lambda_contrastive=0.1),
    "l1": dict(use_mse=False, use_L1=True, use_gdl=False, use_contrastive=False),
    "gdl1": dict(use_mse=False, use_L1=True, use_gdl=True, lambda_gdl=0.5, alpha=1, use_contrastive=True, temperature=0.1,
                 lambda_contrastive=0.05),
}


@pytest.mark.parametrize("tag", list(CASES))
def test_train_step_matches_reference_golden(tag):
    g = load_golden("train_step")
    d, H, Le, Ld, E = (int(v) for v in g["arch"])
    B, S, P = (int(v) for v in g["shape"])
    m, ref = build_pair(d, H, Le, Ld, int(g["seed"]))
    if sd_checksum(ref.state_dict()) != str(g["checksum"]):
        pytest.skip("seeded init differs from the fixture's torch build")
    tr = sdvg_b200.AdamTrainer(m, lr=float(g["lr"]), frames_to_predict=P, **CASES[tag])
    w0 = {k: v.detach().clone() for k, v in ref.named_parameters()}
    for step in range(2):
        batch = OT.make_batch(B, S, E, seed=100 + step).to(DEV)
        losses = tr.step(batch)
        assert abs(float(losses[0]) - float(g[f"{tag}.loss{step}"])) <= 2e-5 * abs(float(g[f"{tag}.loss{step}"])), step
        if step == 0:
            pred = tr.prediction(B, S - 1)
            want = g[f"{tag}.pred0"]
            assert float((pred.cpu() - want).abs().max() / want.abs().max()) < 1e-4
            for k, _ in ref.named_parameters():
                got = tr.gradient(k)
                amax = float(g[f"{tag}.g.{k}.amax"])
                assert float((sample(got).cpu() - g[f"{tag}.g.{k}.s"]).abs().max()) <= TOLG * amax, k
                assert abs(float(got.abs().max()) - amax) <= 2 * TOLG * amax, k
                assert abs(float(got.double().sum()) - float(g[f"{tag}.g.{k}.sum"])) <= TOLG * amax * max(8.0, got.numel() ** 0.5), k
    tr.pull_weights()
    # Adam normalises every element by its own gradient magnitude, so elements whose gradient is tiny against the
    # tensor's maximum amplify rounding differences: bound the error by a fraction of the step size (two steps
    # move an element by up to 2 lr)
    # (e.g. an attention's key bias, whose true gradient is zero: rounding noise becomes +-lr steps)
    lr = float(g["lr"])
    for k, p in m.named_parameters():
        want = g[f"{tag}.w.{k}.s"]
        assert float((want - sample(w0[k])).abs().max()) > 0.5 * lr, k             # the fixture did move this tensor
        err = (sample(p).cpu() - want).abs()
        big = g[f"{tag}.g.{k}.s"].abs() > 1e-2 * float(g[f"{tag}.g.{k}.amax"])
        assert float(err.max()) <= 2.1 * lr, k
        if big.any() and tag == "c5":     # smooth loss only: L1 / GDL(alpha 1) gradients are sign functions
            assert float(err[big].max()) <= 0.5 * lr and float(err[big].mean()) <= 0.05 * lr, k


def test_full_gradients_vs_oracle_and_partial_loss_window():
    """Every element of every gradient against autograd of the oracle; frames_to_predict < S_tgt; B = 1."""
    for (B, S, P, kw) in ((4, 6, 5, CASES["c5"]), (2, 6, 2, CASES["gdl1"]), (1, 4, 3, CASES["l1"])):
        m, ref = build_pair(64, 2, 2, 2, seed=5)
        tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=P, **kw)
        opt = torch.optim.Adam(ref.parameters(), lr=1e-5)
        batch = OT.make_batch(B, S, 256, seed=7)
        loss, pred, grads = OT.train_step_ref(ref, opt, batch, P, **kw)
        losses = tr.step(batch.to(DEV))
        assert abs(float(losses[0]) - float(loss)) <= 2e-5 * abs(float(loss))
        for k, gr in grads.items():
            got = tr.gradient(k).cpu()
            assert float((got - gr).abs().max()) <= TOLG * float(gr.abs().max()) + 1e-12, (B, S, P, k)
        tr.pull_weights()
        for k, p in ref.named_parameters():       # first Adam step = lr * sign(g) wherever |g| >> eps: same sign, same step
            err = (m.state_dict()[k].cpu() - p.detach()).abs()
            big = grads[k].abs() > 1e-2 * grads[k].abs().max()
            assert float(err.max()) <= 2.1e-5, k
            if big.any():
                assert float(err[big].max()) <= 1e-6, k


def test_c5_architecture_step_vs_oracle():
    """BASELINE config 5 (11_19_wallpushups_all_losses_test: d1024 H16 12e/12d, E=1024, batch 16, MSE + GDL(alpha 2) +
    0.1 BiPatchNCE, Adam lr 1e-5) - one step on a reduced batch of 4 clips.  At the far end of a 24-layer backward
    chain the fp32 reference itself is 4e-4..8e-4 away from its float64 run (layer-0 attention, embedding), so the
    yardstick is the float64 oracle: ours must be within 1e-4 plus twice the fp32 reference's own error."""
    c = sdvg_b200.CONFIGS["11_19_wallpushups_all_losses_test"]
    arch = (c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"])
    m, ref = build_pair(*arch, seed=0, frame_size=c["frame_size"])
    torch.manual_seed(0)
    ref64 = RefTransformer(0, *arch, 0.0, frame_size=c["frame_size"]).double()
    tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=5, **CASES["c5"])
    batch = OT.make_batch(4, 6, 1024, seed=9)
    loss, pred, grads = OT.train_step_ref(ref, torch.optim.Adam(ref.parameters(), lr=1e-5), batch, 5, **CASES["c5"])
    _, _, grads64 = OT.train_step_ref(ref64, torch.optim.Adam(ref64.parameters(), lr=1e-5), batch.double(), 5, **CASES["c5"])
    losses = tr.step(batch.to(DEV))
    assert abs(float(losses[0]) - float(loss)) <= 2e-5 * abs(float(loss))
    assert float((tr.prediction(4, 5).cpu() - pred).abs().max() / pred.abs().max()) < 1e-4
    for k, g64 in grads64.items():
        scale = float(g64.abs().max())
        ours = float((tr.gradient(k).cpu().double() - g64).abs().max()) / scale
        ref32 = float((grads[k].double() - g64).abs().max()) / scale
        assert ours <= TOLG + 2.0 * ref32, (k, ours, ref32)


def test_c5_bench_shape_step_with_dropout_vs_same_mask_oracle():
    """The bench shape of BASELINE config 5: 16 clips per GPU, DROPOUT_P 0.1 as in the yaml - training steps against the
    float64 oracle run under the SAME dropout masks (oracle/dropout.py restates the library's counter-based hash).

    Forward (loss, prediction): strict.  Gradients: at this size the reference's own step function is not continuous at
    fp32 resolution, measured with tools/c5_diag.py and tools/loss_kink_exp.py on B200 / CPU:
      * Trainer.criterion's GDL term has sign kinks: perturbing the float64 prediction by 1e-5 of its range (the distance
        between any two fp32 implementations: libsdvg 1.1e-5, torch fp32 5.6e-6) moves 2-12 elements of dL/dpred by
        2-17 % of its largest entry;
      * 4.3 million ReLU evaluations per step put a handful of pre-activations within rounding distance of zero: the unit
        opens in one implementation and not in the other, ONE row of that layer's linear1.weight gradient differs by
        O(1e-2) of the tensor's maximum (seen as a single outlier row, e.g. decoder layer 11 row 1913: 4e-2, every
        other row <= 1.3e-5) and every tensor upstream of it by a dense ~3e-4 (one token of 80 carries a 2 % different
        gradient).
    torch's own fp32 run is 3e-4 ... 3e-2 from its float64 run on these tensors for the same reason, on different rows.
    So: (a) the fused step is held to strict forward values and to gradient bounds that tolerate the loss kinks (whole
    gradient vector within 3e-2); (b) the backward pass is tested on its own with the oracle's dL/dpred as upstream
    gradient: elementwise 1e-4 on every tensor between the loss and the last ReLU, and kink-tolerant bounds elsewhere
    (one flipped row <= 2e-1 of its tensor's maximum, median tensor <= 5e-3, whole vector <= 3e-2) - which still catch
    any systematic error (a wrong mask, tile or scale is O(1) on whole tensors).  Elementwise 1e-4 agreement on every
    tensor is tested where no kink is hit (the other tests of this file)."""
    from oracle import dropout as D
    c = sdvg_b200.CONFIGS["11_19_wallpushups_all_losses_test"]
    arch = (c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"])
    m, ref = build_pair(*arch, seed=0, frame_size=c["frame_size"])
    m.dropout_p = 0.1
    seed = 0x5EED_0C5
    sd64 = {k: v.double() for k, v in ref.state_dict().items()}

    def errors(grad_of, g64):
        per_tensor, num, den = {}, 0.0, 0.0
        for k, gr in g64.items():
            d = grad_of(k).cpu().double() - gr
            per_tensor[k] = float(d.abs().max() / gr.abs().max())
            num += float(d.pow(2).sum()); den += float(gr.pow(2).sum())
        vals = sorted(per_tensor.values())
        return per_tensor, max(vals), vals[len(vals) // 2], (num / den) ** 0.5

    # (a) the fused step (criterion gradient computed by the library from ITS prediction): forward strict, gradients
    # within what the loss kinks allow
    tr = sdvg_b200.AdamTrainer(m, lr=0.0, frames_to_predict=5, seed=seed, **CASES["c5"])     # lr = 0: the weights stay put
    assert tr.dropout == 0.1
    batch = OT.make_batch(16, 6, 1024, seed=12)
    loss64, pred64, g64 = OT.train_grads_functional(sd64, arch[1], batch.double(), 5, drop=D.Dropper(0.1, seed, 1), **CASES["c5"])
    losses = tr.step(batch.to(DEV))
    assert abs(float(losses[0]) - float(loss64)) <= 2e-5 * abs(float(loss64))
    assert float((tr.prediction(16, 5).cpu().double() - pred64).abs().max() / pred64.abs().max()) < 1e-4
    _, worst, median, whole = errors(tr.gradient, g64)
    assert worst <= 2e-1 and median <= 1e-2 and whole <= 3e-2, (worst, median, whole)

    # (b) the backward pass on its own: the next batch through the autograd bridge (trainers/trainer.py:141,164 -
    # pred = model(...); loss.backward()), with the float64 oracle's dL/dpred as the upstream gradient, so the criterion's
    # kinks are out of the comparison.  Tensors between the loss and the last ReLU of the model: strict.  The rest carries
    # the ReLU flips of 24 layers (tools/c5_diag2.py: a few rows per linear1.weight hold 50-84 % of that tensor's error;
    # the dense remainder accumulates towards the start of the chain - embedding.weight and encoder layer 0's in_proj
    # are 7e-3 in the Frobenius norm and hold 97 % of the whole vector's 4-11e-3): one row <= 2e-1, median tensor <= 5e-3,
    # whole vector <= 3e-2.
    batch = OT.make_batch(16, 6, 1024, seed=14)
    _, pred64, g64, dpred64 = OT.train_grads_functional(sd64, arch[1], batch.double(), 5, drop=D.Dropper(0.1, seed, 2),
                                                         return_dpred=True, **CASES["c5"])
    m.train()
    m.dropout_seed = seed
    x = batch.to(DEV)
    pred = m(x, x[:, :-1].contiguous(), m.get_tgt_mask(5).to(DEV))
    assert float((pred.detach().cpu().double() - pred64).abs().max() / pred64.abs().max()) < 1e-4
    m.zero_grad()
    pred.backward(dpred64.float().to(DEV))
    params = dict(m.named_parameters())
    per_tensor, worst, median, whole = errors(lambda k: params[k].grad, g64)
    after_last_relu = [k for k in g64 if k.startswith(("out.", "transformer.decoder.norm.", "transformer.decoder.layers.11.norm3.",
                                                       "transformer.decoder.layers.11.linear2."))]
    assert len(after_last_relu) == 8
    for k in after_last_relu:
        assert per_tensor[k] <= TOLG, (k, per_tensor[k])
    assert worst <= 2e-1 and median <= 5e-3 and whole <= 3e-2, (worst, median, whole)


def test_odd_widths_and_head_sizes():
    """d = 96 with 4 heads (head dim 24: not a multiple of 32, operand planes padded 96 -> 128 columns, 96-row weight
    tiles) and d = 32 with 4 heads (head dim 8), one encoder / zero... one decoder layer, 3-token windows."""
    for (d, H, Le, Ld, B, S, P) in ((96, 4, 1, 1, 3, 6, 5), (32, 4, 2, 1, 2, 4, 2)):
        m, ref = build_pair(d, H, Le, Ld, seed=12)
        tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=P, **CASES["c5"])
        opt = torch.optim.Adam(ref.parameters(), lr=1e-5)
        batch = OT.make_batch(B, S, 256, seed=13)
        loss, pred, grads = OT.train_step_ref(ref, opt, batch, P, **CASES["c5"])
        losses = tr.step(batch.to(DEV))
        assert abs(float(losses[0]) - float(loss)) <= 2e-5 * abs(float(loss)), d
        for k, gr in grads.items():
            assert float((tr.gradient(k).cpu() - gr).abs().max()) <= TOLG * float(gr.abs().max()) + 1e-12, (d, k)


def test_large_training_batch_takes_the_multi_tile_paths():
    """B = 64 clips (the reference's maximum batch): 384 / 320 token rows -> several 128-row tiles, CTA-pair GEMMs,
    weight-gradient GEMMs with K = 384, column sums over several row chunks; B = 30 for rows just above one tile.

    At this size an instance usually has a value on a kink: with ~2 M FFN pre-activations per pass the smallest |z| is
    ~1e-6 of the rms (relu_margin), and the GDL term has |.| kinks too.  Such a unit can fall on either side in two
    fp32 implementations (measured with seed 15: the fp32 oracle is 6.7e-3 from its own float64 run on out.weight at
    B = 30 while libsdvg is 2e-6 from it; at B = 64 one ReLU of decoder layer 0 opens in libsdvg and not in the oracle)
    and the affected gradient rows then differ by O(1).  That is a discontinuity of the model, not an error of either
    side, so: every instance must be within 2e-2 of the float64 oracle everywhere, and of up to four seeded instances
    one must be within 1e-4 everywhere (no value on a kink)."""
    for B in (64, 30):
        strict_seen = False
        for seed in range(15, 19):
            m, ref = build_pair(64, 2, 1, 2, seed=14)
            ref = ref.double()
            tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=5, **CASES["c5"])
            batch = OT.make_batch(B, 6, 256, seed=seed)
            assert relu_margin(ref, batch.double()) < 1e-4
            loss, pred, grads = OT.train_step_ref(ref, torch.optim.Adam(ref.parameters(), lr=1e-5), batch.double(), 5, **CASES["c5"])
            losses = tr.step(batch.to(DEV))
            assert abs(float(losses[0]) - float(loss)) <= 2e-5 * abs(float(loss))
            assert float((tr.prediction(B, 5).cpu().double() - pred).abs().max() / pred.abs().max()) < 1e-4
            errs = [float((tr.gradient(k).cpu().double() - gr).abs().max() / gr.abs().max()) for k, gr in grads.items()]
            assert max(errs) <= 2e-2, (B, seed, max(errs))
            if max(errs) <= TOLG:
                strict_seen = True
                break
        assert strict_seen, B


@pytest.mark.parametrize("p", [0.0, 0.2])
def test_long_training_windows(p):
    """12 latents per clip (S_src 12 / S_tgt 11: the 16-token instantiations of the attention kernels), loss on the
    last 7 positions, with and without dropout (same-mask oracle)."""
    from oracle import dropout as D
    m, ref = build_pair(64, 2, 1, 1, seed=16)
    tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=7, dropout=p, seed=99, **CASES["c5"])
    batch = OT.make_batch(3, 12, 256, seed=17)
    sd64 = {k: v.double() for k, v in ref.state_dict().items()}
    loss, pred, grads = OT.train_grads_functional(sd64, 2, batch.double(), 7, drop=D.Dropper(p, 99, 1) if p else None, **CASES["c5"])
    losses = tr.step(batch.to(DEV))
    assert abs(float(losses[0]) - float(loss)) <= 2e-5 * abs(float(loss))
    for k, gr in grads.items():
        assert float((tr.gradient(k).cpu().double() - gr).abs().max()) <= TOLG * float(gr.abs().max()) + 1e-12, k
    with pytest.raises(RuntimeError):
        tr2 = sdvg_b200.AdamTrainer(build_pair(64, 2, 1, 1, seed=16)[0], frames_to_predict=5, **CASES["c5"])
        tr2.step(OT.make_batch(2, 18, 256, seed=1).to(DEV))          # 18 tokens > 16


def test_trained_weights_reach_state_dict_and_inference():
    """After AdamTrainer steps, model.state_dict() (torch.save at trainers/trainer.py:294) returns the trained values
    without an explicit pull, the eval-mode forward uses them, and growing the batch after training started raises
    instead of silently dropping the Adam moments."""
    m, ref = build_pair(64, 2, 1, 2, seed=9)
    _, ref0 = build_pair(64, 2, 1, 2, seed=9)                       # stays at the initial weights
    tr = sdvg_b200.AdamTrainer(m, lr=1e-2, frames_to_predict=5, **CASES["c5"])
    opt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    batch = OT.make_batch(4, 6, 256, seed=2)
    before = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    OT.train_step_ref(ref, opt, batch, 5, **CASES["c5"])
    tr.step(batch.to(DEV))
    sd = m.state_dict()                                             # no pull_weights() call
    moved = max(float((sd[k].cpu() - before[k]).abs().max()) for k in before)
    assert moved > 5e-3
    big = [k for k, p in ref.named_parameters()]
    assert max(float((sd[k].cpu() - ref.state_dict()[k]).abs().max()) for k in big) <= 2.1e-2   # same +-lr steps
    x = batch[:, :5].contiguous().to(DEV)
    ref.eval(); ref0.eval(); m.eval()
    with torch.no_grad():
        want = ref(x.cpu(), x.cpu(), ref.get_tgt_mask(5))
        stale = ref0(x.cpu(), x.cpu(), ref.get_tgt_mask(5))
        got = m(x, x, "causal").cpu()
    assert float((got - want).abs().max()) < 0.25 * float((got - stale).abs().max())   # trained, not the initial, weights
    with pytest.raises(RuntimeError, match="rebuilt"):
        tr.step(OT.make_batch(m._limits["max_clips"] + 8, 6, 256, seed=3).to(DEV))


def test_data_parallel_shards_equal_global_batch():
    """Two replicas, each with half of the clips and pe_index = global positions: the mean of their gradients is the
    gradient of the global batch (what the NCCL all-reduce + 1/world Adam step computes), SURVEY.md 8(e)."""
    kw = CASES["c5"]
    B, S = 6, 6
    batch = OT.make_batch(B, S, 256, seed=21)
    m, ref = build_pair(64, 2, 1, 2, seed=6)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-5)
    _, _, grads = OT.train_step_ref(ref, opt, batch, 5, **kw)
    shard_grads = []
    for r in range(2):
        mr, _ = build_pair(64, 2, 1, 2, seed=6)
        tr = sdvg_b200.AdamTrainer(mr, lr=1e-5, frames_to_predict=5, **kw)
        lo, hi = r * B // 2, (r + 1) * B // 2
        tr.step(batch[lo:hi].to(DEV), pe_index=torch.arange(lo, hi))
        shard_grads.append({k: tr.gradient(k).clone() for k in grads})
    for k, gr in grads.items():
        mean = 0.5 * (shard_grads[0][k] + shard_grads[1][k]).cpu()
        assert float((mean - gr).abs().max()) <= TOLG * float(gr.abs().max()), k


@pytest.mark.parametrize("precision,tol", [("fp16", 2e-2), ("bf16", 1e-1)])
def test_other_16bit_training_modes_run(precision, tol):
    """fp16 / bf16 operand planes everywhere (bf16 has 8 significant bits: loose bound, it is not a parity mode)."""
    m, ref = build_pair(64, 2, 1, 1, seed=8, precision=precision)
    tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=5, **CASES["c5"])
    batch = OT.make_batch(3, 6, 256, seed=3)
    loss, _, grads = OT.train_step_ref(ref, torch.optim.Adam(ref.parameters(), lr=1e-5), batch, 5, **CASES["c5"])
    losses = tr.step(batch.to(DEV))
    assert abs(float(losses[0]) - float(loss)) <= tol * abs(float(loss))
    for k, gr in grads.items():
        assert float((tr.gradient(k).cpu() - gr).norm()) <= 5 * tol * float(gr.norm()) + 1e-12, k


def test_16bit_training_mode_and_errors():
    """`mixed` precision (single fp16 operand planes beyond the first layers): gradients of the smooth config-5 loss
    within 2e-2 of the reference in the Frobenius norm per tensor.  (Max-norm is not meaningful in 16-bit mode: a
    pre-activation within rounding distance of zero flips its ReLU gate and changes single gradient rows by O(1/M);
    likewise an L1 / GDL(1) loss has sign-function gradients.)"""
    m, ref = build_pair(64, 2, 1, 1, seed=8, precision="mixed")
    tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=5, **CASES["c5"])
    opt = torch.optim.Adam(ref.parameters(), lr=1e-5)
    batch = OT.make_batch(3, 6, 256, seed=3)
    loss, _, grads = OT.train_step_ref(ref, opt, batch, 5, **CASES["c5"])
    losses = tr.step(batch.to(DEV))
    assert abs(float(losses[0]) - float(loss)) <= 5e-3 * abs(float(loss))
    for k, gr in grads.items():
        assert float((tr.gradient(k).cpu() - gr).norm()) <= 2e-2 * float(gr.norm()) + 1e-12, k
    with pytest.raises(RuntimeError):
        sdvg_b200.AdamTrainer(m, use_mse=True, use_L1=True)
    with pytest.raises(RuntimeError):
        tr.step(batch)                                                # CPU tensor
    with pytest.raises(RuntimeError):
        sdvg_b200.AdamTrainer(m, dropout=1.0)


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_training_matches_oracle_under_the_same_masks(p):
    """model.train() with DROPOUT_P > 0: the library draws its masks from a counter-based hash (include/sdvg.h,
    sdvg_train_set_dropout); oracle/dropout.py restates that hash, so the oracle runs nn.Transformer's dropout sites
    (embedding + PE, attention probabilities, sub-layer outputs, FFN hidden) under exactly the same masks and autograd
    gives the reference gradients.  Two steps: the masks change with the step counter, the weights with Adam.
    Yardstick = the float64 run of the oracle (at p = 0.5 the fp32 oracle itself is 1e-4 from it on layer-0 tensors)."""
    from oracle import dropout as D
    seed = 0x1234_5678_9ABC
    m, ref = build_pair(64, 2, 2, 2, seed=5)
    m.dropout_p = p                                   # what Transformer(..., dropout_p=p) would carry
    tr = sdvg_b200.AdamTrainer(m, lr=1e-3, frames_to_predict=5, seed=seed, **CASES["c5"])
    assert tr.dropout == p
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    losses_seen = []
    for step in (1, 2):
        batch = OT.make_batch(4, 6, 256, seed=30 + step)
        sd64 = {k: v.double() for k, v in sd.items()}
        loss, pred, grads = OT.train_grads_functional(sd64, 2, batch.double(), 5, drop=D.Dropper(p, seed, step), **CASES["c5"])
        losses = tr.step(batch.to(DEV))
        assert abs(float(losses[0]) - float(loss)) <= 2e-5 * abs(float(loss)), step
        assert float((tr.prediction(4, 5).cpu().double() - pred).abs().max() / pred.abs().max()) < 1e-4, step
        for k, gr in grads.items():
            assert float((tr.gradient(k).cpu().double() - gr).abs().max()) <= TOLG * float(gr.abs().max()) + 1e-12, (step, k)
        losses_seen.append(float(loss))
        tr.pull_weights()
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    # without dropout the same first batch gives a different loss (the masks really were applied)
    l0, _, _ = OT.train_grads_functional(ref.state_dict(), 2, OT.make_batch(4, 6, 256, seed=31), 5, **CASES["c5"])
    assert abs(float(l0) - losses_seen[0]) > 1e-3 * abs(float(l0))


@pytest.mark.parametrize("tag", ["c5", "gdl1"])
def test_reference_loop_body_runs_unmodified(tag):
    """The literal body of Trainer.train_loop (trainers/trainer.py:126-165) with only the model import swapped:
    ``pred = model(new_batch, y_input, tgt_mask)``, the reference's own torch criterion, ``opt.zero_grad();
    loss.backward(); opt.step()`` with ``torch.optim.Adam(model.parameters())`` (:365) - against the golden values of the
    unmodified reference (losses, gradient samples, weights after two steps)."""
    from oracle import losses as OL
    g = load_golden("train_step")
    d, H, Le, Ld, E = (int(v) for v in g["arch"])
    B, S, P = (int(v) for v in g["shape"])
    model, ref = build_pair(d, H, Le, Ld, int(g["seed"]))
    if sd_checksum(ref.state_dict()) != str(g["checksum"]):
        pytest.skip("seeded init differs from the fixture's torch build")
    lr = float(g["lr"])
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    loss_fn = OL.criterion(**CASES[tag])
    model.train()
    for step in range(2):
        new_batch = OT.make_batch(B, S, E, seed=100 + step).to(DEV)
        y_input = new_batch[:, :-1]
        y_expected = new_batch[:, 1:]
        y_expected = y_expected.permute(1, 0, 2)
        tgt_mask = model.get_tgt_mask(y_input.size(1)).to(DEV)
        pred = model(new_batch, y_input, tgt_mask)
        loss = loss_fn(pred[-P:], y_expected[-P:])
        opt.zero_grad()
        loss.backward()
        opt.step()
        want = float(g[f"{tag}.loss{step}"])
        assert abs(float(loss) - want) <= 2e-5 * abs(want), step
        if step == 0:
            assert float((pred.detach().cpu() - g[f"{tag}.pred0"]).abs().max() / g[f"{tag}.pred0"].abs().max()) < 1e-4
            for k, p in model.named_parameters():
                amax = float(g[f"{tag}.g.{k}.amax"])
                assert float((sample(p.grad).cpu() - g[f"{tag}.g.{k}.s"]).abs().max()) <= TOLG * amax, k
    for k, p in model.named_parameters():
        err = (sample(p).cpu() - g[f"{tag}.w.{k}.s"]).abs()
        assert float(err.max()) <= 2.1 * lr, k
    # validation afterwards (model.eval(), no_grad) sees the weights torch's optimiser wrote
    model.eval()
    with torch.no_grad():
        out = model(new_batch, y_input.contiguous(), "causal")
    ref.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()})
    ref.eval()
    with torch.no_grad():
        want = ref(new_batch.cpu(), y_input.cpu(), ref.get_tgt_mask(S - 1))
    assert float((out.cpu() - want).abs().max() / want.abs().max()) < 1e-4


def test_checkpoint_files_round_trip_with_the_reference_module(tmp_path):
    """prediction/predict.py:51 and trainers/trainer.py:472-479: a state_dict FILE written by the reference module loads
    into the drop-in, is trained one step, saved with torch.save(model.state_dict()) and loads back into the reference."""
    torch.manual_seed(3)
    ref = RefTransformer(0, 64, 2, 2, 2, 0.0, frame_size=64)
    f_in, f_out = tmp_path / "ref_ckpt.pt", tmp_path / "ours_ckpt.pt"
    torch.save(ref.state_dict(), f_in)
    m = sdvg_b200.Transformer(0, 64, 2, 2, 2, 0.0, frame_size=64, precision="fp32")
    m.load_state_dict(torch.load(f_in))
    m = m.to(DEV)
    batch = OT.make_batch(4, 6, 256, seed=21)
    tr = sdvg_b200.AdamTrainer(m, lr=1e-3, frames_to_predict=5, **CASES["c5"])
    tr.step(batch.to(DEV))
    torch.save(m.state_dict(), f_out)
    back = RefTransformer(0, 64, 2, 2, 2, 0.0, frame_size=64)
    sd = torch.load(f_out, map_location="cpu")
    assert sorted(sd) == sorted(ref.state_dict())
    back.load_state_dict(sd)                                                   # strict: same keys, same shapes
    moved = max(float((sd[k] - v).abs().max()) for k, v in ref.state_dict().items() if "pos_encoding" not in k)
    assert moved > 1e-4                                                        # the file holds the TRAINED weights
    back.eval()
    m.eval()
    x = batch[:, :5].contiguous()
    with torch.no_grad():
        want = back(x, x, back.get_tgt_mask(5))
        got = m(x.to(DEV), x.to(DEV), "causal").cpu()
    assert float((got - want).abs().max() / want.abs().max()) < 1e-4


def test_trained_weights_survive_an_engine_rebuild():
    """After AdamTrainer.step the trained weights live only in the engine.  Growing a workspace limit (a rollout with a
    longer history), changing the precision or freeing the handle rebuilds the engine: the weights must be rescued
    first, not replaced by the stale module parameters."""
    twins = [build_pair(64, 2, 2, 2, seed=5)[0] for _ in range(2)]
    batch = OT.make_batch(4, 6, 256, seed=7).to(DEV)
    for m in twins:
        sdvg_b200.AdamTrainer(m, lr=1e-3, frames_to_predict=5, **CASES["c5"]).step(batch)
    want = {k: v.detach().clone() for k, v in twins[1].state_dict().items()}      # plain pull
    m = twins[0]
    before = {k: v.detach().clone() for k, v in torch.nn.Module.state_dict(m).items()}   # stale parameters, no pull
    m.eval()
    ctx = torch.randn(3, 40, 256, generator=torch.Generator().manual_seed(1)).to(DEV)
    out = sdvg_b200.rollout(m, ctx, 2, 5)                   # history 42 > the default 32: reserve() -> engine rebuilt
    got = m.state_dict()
    assert any(not torch.equal(before[k], want[k]) for k in want)                # training did move the weights
    for k in want:
        assert torch.equal(got[k], want[k]), k
    twins[1].eval()
    assert torch.equal(out, sdvg_b200.rollout(twins[1], ctx, 2, 5))             # and the rebuilt engine runs on them
    m.set_precision("mixed")                                                      # another rebuild: still the trained weights
    for k in want:
        assert torch.equal(m.state_dict()[k], want[k]), k
