"""Recipe for ``baseline/_ref``: the UNMODIFIED reference model, so that the reference arm of bench.py
(``--impl reference`` and the ``cpu_baseline`` leg) times the reference's own code on the GPU box's host cores.

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python on top of
``torch.nn.Transformer``: there is nothing to compile, its "build" is making the three modules of the inference path and
the yaml files of the five BASELINE.json configs importable next to the repo.  They are copied from where they lie under
/root/reference into the git-ignored ``baseline/_ref/`` (never into the history; the directory is not gpurun-ignored, so
it travels to the GPU box like the built ``libsdvg.so``).  /root/reference does not exist on the GPU box: there this
script is a no-op and ``load_reference()`` uses what was installed in the build container, or returns None (bench.py then
falls back to the oracle port, oracle/ref_module.py, and says ``kind: "port"``).

    python oracle/install_ref.py        # also run by __graft_entry__.build()

What is copied (reference path : why):
    models/transformer.py           the module under test
    models/positional_encoding.py   imported by it (row a5)
    utils/config.py                 its constructor parses argv + ./config/<name>.yml (models/transformer.py:23,28-29)
    config/<the five BASELINE configs>.yml
prediction/predict.py is NOT usable: it imports diffusers (absent from this image) at module top and loads a VAE; its
27-line ``predict`` and the rollout loop are driven through oracle/rollout.py, which calls the reference MODEL object.
"""
import contextlib
import os
import shutil
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/transformer.py", "models/positional_encoding.py", "utils/config.py"]
CONFIGS = ["1_17_ball_complex_L1_64", "1_15_kitti_L1_64", "11_27_ucf_final", "11_20_wallpushups_dim_2048",
           "11_19_wallpushups_all_losses_test"]


def install():
    """Copy the files; returns the destination, or None when /root/reference is not there (GPU box)."""
    if not os.path.isdir(REF):
        return None
    files = FILES + [f"config/{c}.yml" for c in CONFIGS if os.path.exists(os.path.join(REF, "config", c + ".yml"))]
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    return DST


def available():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


@contextlib.contextmanager
def _reference_process_state(config_name):
    """The reference constructor reads sys.argv and ./config/<name>.yml relative to the cwd."""
    argv, cwd = sys.argv, os.getcwd()
    sys.argv = ["x", "--dataset", "ball", "--config", config_name]
    sys.path.insert(0, DST)
    os.chdir(DST)
    try:
        yield
    finally:
        os.chdir(cwd)
        sys.argv = argv
        sys.path.remove(DST)


def load_reference(config_name, arch, seed=0):
    """The reference's own ``Transformer`` for a BASELINE config, constructed by the reference's constructor with the
    reference's argument order (prediction/predict.py:47-49), or None when baseline/_ref is not installed / the yaml of
    that config is missing.  ``arch`` = (dim_model, num_heads, num_encoder_layers, num_decoder_layers, dropout_p)."""
    if not available() or not os.path.exists(os.path.join(DST, "config", config_name + ".yml")):
        return None
    import torch
    with _reference_process_state(config_name):
        from models.transformer import Transformer      # noqa: the reference, unmodified
        torch.manual_seed(seed)
        return Transformer(0, *arch).eval()


if __name__ == "__main__":
    print("installed" if install() else "no /root/reference here", DST, "available:", available())
