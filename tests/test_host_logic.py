"""CPU: host-side mirror of the reference interface (module contract, state_dict, masks, sharding)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

import sdvg_b200
from conftest import ROOT, load_golden, ref_model_from_golden
from oracle import rollout as R
from oracle.ref_module import RefTransformer


def test_state_dict_matches_reference_keys_and_init():
    g = load_golden("tiny_forward")
    ref_keys = sorted(k[3:] for k in g if k.startswith("sd."))
    torch.manual_seed(0)
    ours = sdvg_b200.Transformer(0, 32, 4, 1, 1, 0.1, frame_size=64)
    sd = ours.state_dict()
    assert sorted(sd) == ref_keys
    for k in ref_keys:
        assert tuple(sd[k].shape) == tuple(g["sd." + k].shape), k
        # same construction order as the reference -> identical seeded init
        assert torch.equal(sd[k], g["sd." + k]), k
    ours.load_state_dict({k: g["sd." + k] for k in ref_keys})          # reference checkpoints load


def test_default_ctor_shapes_like_reference():
    m = sdvg_b200.Transformer(frame_size=64)                            # Transformer() defaults, models/transformer.py:14-19
    assert m.dim_model == 256 and m.height == 64 and m.width == 64 and m.compression == 8
    assert len(m.transformer.encoder.layers) == 6 and len(m.transformer.decoder.layers) == 6
    assert m.embedding.in_features == 256 and m.out.out_features == 256
    assert m.transformer.encoder.layers[0].linear1.out_features == 2048
    assert tuple(m.positional_encoder.pos_encoding.shape) == (64, 1, 256)


def test_ctor_reads_frame_size_from_argv_and_yaml(tmp_path, monkeypatch):
    """models/transformer.py:23,28-29: without frame_size the ctor parses --dataset/--config and ./config/<name>.yml."""
    (tmp_path / "config").mkdir()
    (tmp_path / "config" / "demo.yml").write_text("FRAME_SIZE: 128\nDIM_MODEL:\n - 64\n")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["x", "--dataset", "ball", "--config", "demo"])
    m = sdvg_b200.Transformer(0, 32, 4, 1, 1, 0.1)
    assert m.height == 128 and m.embedding.in_features == 1024 and m.config.CONFIG_NAME == "demo"
    monkeypatch.setattr(sys, "argv", ["x"])
    with pytest.raises(SystemExit):                                      # argparse: required flags, like the reference
        sdvg_b200.Transformer(0, 32, 4, 1, 1, 0.1)


def test_masks_match_reference():
    m = sdvg_b200.Transformer(0, 32, 4, 1, 1, 0.1, frame_size=64)
    r = RefTransformer(0, 32, 4, 1, 1, 0.1, frame_size=64)
    for s in (1, 2, 5, 6, 10):
        assert torch.equal(m.get_tgt_mask(s), r.get_tgt_mask(s))
    x = torch.tensor([1, 2, 3, 0, 0, 0])
    assert torch.equal(m.create_pad_mask(x, 0), r.create_pad_mask(x, 0))


def test_b65_raises_before_any_device_work():
    m = sdvg_b200.Transformer(0, 32, 4, 1, 1, 0.1, frame_size=64).eval()
    x = torch.zeros(65, 2, 256)
    with pytest.raises(RuntimeError, match=r"size of tensor a \(65\)"):
        m(x, x)
    with pytest.raises(RuntimeError, match="padding masks"):
        m(x[:2], x[:2], None, torch.zeros(2, 2, dtype=torch.bool))
    m.train()
    # train() mode is the autograd bridge of trainers/trainer.py:141: causal mask only, CUDA only, rollouts need eval()
    with pytest.raises(RuntimeError, match="causal target mask"):
        m(x[:2], x[:2])
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(x[:2], x[:2], m.get_tgt_mask(2))
    with pytest.raises(RuntimeError, match="inference path"):
        m.rollout(x[:2], 2)


def test_configs_table():
    c = sdvg_b200.CONFIGS
    assert c["1_17_ball_complex_L1_64"]["dim_model"] == 2048 and c["1_17_ball_complex_L1_64"]["num_decoder_layers"] == 8
    assert c["11_20_wallpushups_dim_2048"]["num_encoder_layers"] == 6
    assert sdvg_b200.latent_dim(64) == 256 and sdvg_b200.latent_dim(128) == 1024


def test_shard_bounds_cover_and_align():
    for B in (1, 8, 63, 64, 65, 128, 1000, 1024, 8192):
        for W in (1, 2, 3, 4, 8):
            spans = [sdvg_b200.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            for a, b in spans:
                assert a % 64 == 0 or a == B
    assert sdvg_b200.pe_index_for(60, 70).tolist() == [60, 61, 62, 63, 0, 1, 2, 3, 4, 5]


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    g = load_golden("small_rollout")
    model = ref_model_from_golden(g)
    sd = model.state_dict()
    ctx = torch.randn(130, 6, 256, generator=torch.Generator().manual_seed(5))    # > 2 chunks of 64, ragged tail

    def oracle_rollout(c, n_pred, window, pe_index=None):
        from oracle import functional as F

        class M:
            get_tgt_mask = staticmethod(F.causal_mask)

            def __call__(self, s, t, mask):
                return F.forward(sd, s, t, 8, mask, pe_index=pe_index.long())
        return R.rollout_ref(M(), c, n_pred, window)

    with torch.no_grad():
        got = sdvg_b200.rollout_sharded(oracle_rollout, ctx, 2, 5, rank=rank, world_size=world)
        if rank == 0:
            want = R.chunked(lambda c: R.rollout_ref(model, c, 2, 5), ctx)          # the reference in chunks of 64
            torch.save((got, want), tmp)
    dist.destroy_process_group()


def test_sharded_rollout_equals_chunked_reference_gloo(tmp_path):
    """world_size-2 gloo: shards keep PE[i mod 64], gather restores clip order; equals the reference run in
    chunks of <= 64 clips."""
    tmp = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, 29541, tmp), nprocs=2, join=True)
    got, want = torch.load(tmp)
    assert got.shape == want.shape == (130, 2, 256)
    assert R.max_rel_per_frame(got, want).max() < 5e-5


def test_sibling_modules_on_cpu():
    m = sdvg_b200.TransformerFuture(0, 32, 4, 1, 1, 0.1, frame_size=64, frames_to_predict=5)
    keys = set(m.state_dict())
    base = set(sdvg_b200.Transformer(0, 32, 4, 1, 1, 0.1, frame_size=64).state_dict())
    assert keys == base | {"learned_tgt"} and tuple(m.learned_tgt.shape) == (1, 5, 256)
    ident = sdvg_b200.Identity()
    x = torch.arange(24.0).view(2, 3, 4)
    assert torch.equal(ident(x, x), x[:, -1:])                       # models/identity.py:13-16
    assert torch.equal(ident.get_tgt_mask(4), m.get_tgt_mask(4))


def _dp_worker(rank, world, port, tmp):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import functional as F
    from oracle import losses as L
    from oracle.ref_module import RefTransformer
    from oracle.train import make_batch
    torch.manual_seed(4)
    ref = RefTransformer(0, 32, 4, 1, 1, 0.0, frame_size=64)
    names = [k for k, _ in ref.named_parameters()]
    sd = {k: v.detach().clone().requires_grad_(k in names) for k, v in ref.state_dict().items()}
    kw = dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07,
              lambda_contrastive=0.1)
    batch = make_batch(6, 6, 256, seed=12)

    def grads_of(x, pe):
        for k in names:
            sd[k].grad = None
        pred = F.forward(sd, x, x[:, :-1], 4, F.causal_mask(5), pe_index=pe)
        L.criterion(**kw)(pred, x[:, 1:].permute(1, 0, 2)).backward()
        return torch.cat([sd[k].grad.reshape(-1) for k in names])

    per = 6 // world
    flat = grads_of(batch[rank * per:(rank + 1) * per], torch.arange(rank * per, (rank + 1) * per))   # this rank's shard
    split = flat.numel() // 3
    for w in sdvg_b200.allreduce_buckets(flat, split, world):
        w.wait()
    flat /= world                                               # what sdvg_train_adam_step's grad_mul = 1/world applies
    if rank == 0:
        torch.save((flat, grads_of(batch, torch.arange(6))), tmp)
    dist.destroy_process_group()


def test_data_parallel_gradient_buckets_gloo(tmp_path):
    """world_size-2 gloo: the two-bucket sum all-reduce of the flat gradient vector followed by 1/world equals the
    gradient of the global batch (equal shards, mean-reduced losses, pe_index = global clip positions) - the host
    logic of AdamTrainer.step, SURVEY.md section 8e."""
    tmp = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(2, 29543, tmp), nprocs=2, join=True)
    got, want = torch.load(tmp)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())


def test_dropout_hash_restatement_matches_library_header(tmp_path):
    """oracle/dropout.py against the library's own drop_hash / drop_site_key (csrc/common.cuh, __host__ __device__):
    a small host program is compiled with nvcc from the header and run on the CPU (no GPU involved)."""
    import shutil
    import subprocess
    from oracle import dropout as D
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "h.cu"
    src.write_text('#include <cstdio>\n#include "common.cuh"\nint main() {\n'
                   '  const unsigned long long seeds[3] = {0ull, 0x123456789ABCull, 0xFFFFFFFFFFFFFFFFull};\n'
                   '  for (int s = 0; s < 3; ++s) for (unsigned step = 1; step < 4; ++step) for (unsigned site : {0u, 1u, 1004u, 2115u}) {\n'
                   '    const unsigned key = sdvg::drop_site_key(seeds[s], step, site);\n'
                   '    printf("%u", key);\n'
                   '    for (unsigned idx : {0u, 1u, 12345u, 4000000000u}) printf(" %u", sdvg::drop_hash(key, idx));\n'
                   '    printf("\\n");\n  }\n  return 0;\n}\n')
    exe = tmp_path / "h"
    inc = os.path.join(ROOT, "sd-video-gen_b200", "csrc")
    subprocess.run([nvcc, "-std=c++17", "-I", inc, "-o", str(exe), str(src)], check=True, capture_output=True, timeout=300)
    lines = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=60).stdout.strip().splitlines()
    it = iter(lines)
    for seed in (0, 0x123456789ABC, 0xFFFFFFFFFFFFFFFF):
        for step in (1, 2, 3):
            for site in (0, 1, 1004, 2115):
                got = [int(v) for v in next(it).split()]
                key = D.site_key(seed, step, site)
                want = [key] + [int(D.fmix32((idx * 0x9E3779B1 + key) & D.M32)) for idx in (0, 1, 12345, 4000000000)]
                assert got == want, (seed, step, site)


def test_latent_cache_reads_the_reference_npy_layout(tmp_path):
    """utils/preprocess.py:27-33 stores one (1, 4, h, w) float32 array per frame next to the png; a clip is the sorted
    file list and E = 4 h w in C order (utils/sd_utils.py:147-149)."""
    import numpy as np
    g = torch.Generator().manual_seed(4)
    clips = torch.randn(3, 7, 4, 8, 8, generator=g)
    dirs = []
    for b in range(3):
        d = tmp_path / f"clip{b}"
        d.mkdir()
        for t in range(7):
            np.save(d / f"frame_{t:03d}.npy", clips[b, t][None].numpy())          # what preprocess.py writes
            (d / f"frame_{t:03d}.png").write_bytes(b"")                            # the images sit beside them
        dirs.append(str(d))
    x = sdvg_b200.load_latent_clips(dirs, frames=5)
    assert tuple(x.shape) == (3, 5, 256) and x.dtype == torch.float32
    assert torch.equal(x, clips[:, :5].reshape(3, 5, -1))
    xs = sdvg_b200.load_latent_clips(dirs, frames=5, use_sos=True)
    assert tuple(xs.shape) == (3, 6, 256) and bool((xs[:, 0] == sdvg_b200.SOS_VALUE).all()) and torch.equal(xs[:, 1:], x)
    out_dir = tmp_path / "pred"
    paths = sdvg_b200.save_latent_frames(str(out_dir), x[0])
    assert np.load(paths[2]).shape == (1, 4, 8, 8)
    assert torch.equal(sdvg_b200.load_latent_frames(paths), x[0])
    with pytest.raises(ValueError):
        sdvg_b200.load_latent_clips(dirs, frames=9)
