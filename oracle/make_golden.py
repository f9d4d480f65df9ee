"""Generate tests/golden/*.npz by running the UNMODIFIED reference module.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only
(/root/reference does not exist on the GPU box):

    python oracle/make_golden.py            # writes tests/golden/*.npz

Recipe: the reference constructor parses sys.argv and reads
./config/<name>.yml relative to the cwd (models/transformer.py:23, utils/config.py:10),
so we chdir into the reference and fake argv.  prediction/predict.py cannot be imported
(it pulls diffusers at module top, predict.py:7) - its 27-line ``predict`` and loop are
driven through oracle/rollout.py, which calls the reference *model* object.

Every case stores inputs, the reference's outputs and either the full state_dict
(tiny models) or the construction seed plus a per-tensor float64 checksum (bigger
models, re-created with oracle.ref_module.RefTransformer which consumes the RNG in
the reference's order; the checksum catches any divergence).
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, ".."))


def import_reference(config_name):
    sys.argv = ["x", "--dataset", "ball", "--config", config_name]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    os.chdir(REF)
    from models.transformer import Transformer  # noqa: the reference, unmodified
    return Transformer


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build(Transformer, seed, **arch):
    torch.manual_seed(seed)
    m = Transformer(0, arch["d"], arch["H"], arch["Le"], arch["Ld"], 0.1)
    return m.eval()


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print(f"wrote {path}  {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    from oracle import rollout as R
    os.makedirs(OUT, exist_ok=True)
    T64 = import_reference("1_17_ball_complex_L1_64")       # FRAME_SIZE 64 -> E = 256
    torch.set_num_threads(8)
    g = torch.Generator().manual_seed(1234)

    # ---- case 1: tiny forward, full weights stored; S_src=6, S_tgt=5 (the trainer's shapes,
    #      trainers/trainer.py:141) with causal mask, plus an unmasked call and src==tgt.
    arch = dict(d=32, H=4, Le=1, Ld=1)
    m = build(T64, 0, **arch)
    src = torch.randn(3, 6, 256, generator=g)
    tgt = src[:, :-1].contiguous()
    with torch.no_grad():
        o_causal = m(src, tgt, m.get_tgt_mask(5))
        o_nomask = m(src, tgt)
        o_same = m(src, src, m.get_tgt_mask(6))
        # PE-by-batch-index quirk: B=64 works, B=65 raises
        x64 = torch.randn(64, 2, 256, generator=g)
        o_b64 = m(x64, x64, m.get_tgt_mask(2))
        try:
            x65 = torch.randn(65, 2, 256, generator=g)
            m(x65, x65, m.get_tgt_mask(2))
            b65 = "ok"
        except RuntimeError as e:
            b65 = "RuntimeError: " + str(e)[:80]
    save("tiny_forward", arch=np.array([32, 4, 1, 1, 256]), src=src, tgt=tgt, out_causal=o_causal,
         out_nomask=o_nomask, out_same=o_same, x64=x64, out_b64=o_b64, b65=np.array(b65),
         **{"sd." + k: v for k, v in m.state_dict().items()})

    # ---- case 2: small rollout (weights by seed), sliding window + teacher forcing + predict.py-faithful
    arch = dict(d=64, H=8, Le=2, Ld=3)
    m = build(T64, 7, **arch)
    ctx = torch.randn(4, 10, 256, generator=g)
    with torch.no_grad():
        free5 = R.rollout_ref(m, ctx, 4, 5)
        free10 = R.rollout_ref(m, ctx, 3, 10)
        frames = torch.randn(2, 5, 256, generator=g) * R.LATENT_SCALE
        faithful = R.rollout_faithful(m, frames, 4)
        one = R.rollout_faithful(m, frames[:1], 2)        # the reference's literal B=1 case
    save("small_rollout", arch=np.array([64, 8, 2, 3, 256]), seed=7, checksum=np.array(sd_checksum(m.state_dict())),
         ctx=ctx, free5=free5, free10=free10, frames=frames, faithful=faithful, faithful_b1=one)

    # ---- case 3: default-arch heads (hd=32), d=256 6e/6d, one forward, weights by seed
    arch = dict(d=256, H=8, Le=6, Ld=6)
    m = build(T64, 3, **arch)
    x = torch.randn(5, 5, 256, generator=g)
    with torch.no_grad():
        o = m(x, x, m.get_tgt_mask(5))
    save("d256_forward", arch=np.array([256, 8, 6, 6, 256]), seed=3, checksum=np.array(sd_checksum(m.state_dict())),
         x=x, out=o)

    # ---- case 4: BASELINE config C1 (1_17_ball_complex_L1_64: d2048 H8 4e/8d E256), full size:
    #      B=8, 10 ctx -> 3 pred, window 5 and one window-10 step; weights by seed 0.
    arch = dict(d=2048, H=8, Le=4, Ld=8)
    m = build(T64, 0, **arch)
    ctx = torch.randn(8, 10, 256, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        r5 = R.rollout_ref(m, ctx, 3, 5)
        r10 = R.rollout_ref(m, ctx, 1, 10)
        md = build(T64, 0, **arch).double()
        r5_64 = R.rollout_ref(md, ctx.double(), 3, 5)
    save("c1_rollout", arch=np.array([2048, 8, 4, 8, 256]), seed=0, checksum=np.array(sd_checksum(m.state_dict())),
         ctx=ctx, free5=r5, free10=r10, free5_fp64=r5_64.float())
    del m, md

    # ---- case 5: E=1024 latents (FRAME_SIZE 128: 11_20_wallpushups_dim_2048 = C4 d2048 H8 6e/6d), one step B=4
    for mod in [k for k in sys.modules if k.startswith(("models", "utils"))]:
        del sys.modules[mod]
    T128 = import_reference("11_20_wallpushups_dim_2048")
    arch = dict(d=2048, H=8, Le=6, Ld=6)
    m = build(T128, 0, **arch)
    ctx = torch.randn(4, 5, 1024, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        r = R.rollout_ref(m, ctx, 2, 5)
    save("c4_rollout", arch=np.array([2048, 8, 6, 6, 1024]), seed=0, checksum=np.array(sd_checksum(m.state_dict())),
         ctx=ctx, free5=r)


if __name__ == "__main__":
    main()
