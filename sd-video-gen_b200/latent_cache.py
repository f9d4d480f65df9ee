"""Reader / writer of the reference's offline latent cache (utils/preprocess.py:27-33): for every ``<frame>.png`` of a
clip directory the reference stores ``<frame>.npy`` = the VAE latents of that frame, shape (1, 4, H/8, W/8) float32,
already multiplied by 0.18215 (:52).  A clip is the sorted list of those files; flattening (4, h, w) in C order gives
the E = 4 h w vector the Transformer sees (utils/sd_utils.py:147-149, ``reshape(B, T, -1)``).

``load_latent_clips`` turns clip directories into the (B, T, E) host tensor ``rollout_from_host`` takes, so the hot
path can be driven from a preprocessed dataset without the VAE."""
import os

import numpy as np
import torch


def _frame_files(clip_dir):
    files = sorted(f for f in os.listdir(clip_dir) if f.endswith(".npy"))
    if not files:
        raise FileNotFoundError(f"no .npy latent frames in {clip_dir}")
    return [os.path.join(clip_dir, f) for f in files]


def load_latent_frames(paths):
    """[path to <frame>.npy, ...] -> (T, E) float32; every file must hold one (1, 4, h, w) (or (4, h, w)) latent."""
    rows = []
    for p in paths:
        a = np.load(p)
        if a.ndim == 4 and a.shape[0] == 1:
            a = a[0]
        if a.ndim != 3 or a.shape[0] != 4:
            raise ValueError(f"{p}: expected latents of shape (1, 4, h, w), got {a.shape}")
        rows.append(np.ascontiguousarray(a, dtype=np.float32).reshape(-1))
    if len({r.size for r in rows}) != 1:
        raise ValueError("latent frames of one clip must have the same size")
    return torch.from_numpy(np.stack(rows))


def load_latent_clips(clip_dirs, frames=None, use_sos=False, sos_value=2.0):
    """Clip directories -> (B, T, E) float32 host tensor.  ``frames`` keeps the first N frames of every clip (clips must
    then have at least N); ``use_sos`` prepends the SOS frame like encode_batch(use_sos=True) (utils/sd_utils.py:151-153)."""
    clips = []
    for d in clip_dirs:
        x = load_latent_frames(_frame_files(d))
        if frames is not None:
            if x.size(0) < frames:
                raise ValueError(f"{d}: {x.size(0)} frames, {frames} requested")
            x = x[:frames]
        clips.append(x)
    if len({tuple(c.shape) for c in clips}) != 1:
        raise ValueError("clips must have the same number of frames and latent size (pass frames=)")
    out = torch.stack(clips)
    if use_sos:
        out = torch.cat([torch.full((out.size(0), 1, out.size(2)), float(sos_value)), out], dim=1)
    return out


def save_latent_frames(clip_dir, latents, names=None):
    """(T, E) latents -> ``<name>.npy`` files of shape (1, 4, h, w), the layout utils/preprocess.py writes."""
    latents = torch.as_tensor(latents).detach().cpu().float()
    T, E = latents.shape
    s = int(round((E // 4) ** 0.5))
    if 4 * s * s != E:
        raise ValueError(f"E = {E} is not 4 * h * h")
    os.makedirs(clip_dir, exist_ok=True)
    paths = []
    for t in range(T):
        p = os.path.join(clip_dir, (names[t] if names else f"{t:05d}") + ".npy")
        np.save(p, latents[t].reshape(1, 4, s, s).numpy())
        paths.append(p)
    return paths
