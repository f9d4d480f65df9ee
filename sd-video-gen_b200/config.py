"""Config plumbing the model constructor depends on (utils/config.py:27-49) and the five BASELINE.json
configurations as data (the yaml files themselves live in the reference tree, which is not shipped)."""
import argparse
import os
from types import SimpleNamespace

# name -> architecture; values read from the reference's yaml files:
#   config/1_17_ball_complex_L1_64.yml:22-34, 1_15_kitti_L1_64.yml:22-34, 11_27_ucf_final.yml:22-34,
#   11_20_wallpushups_dim_2048.yml:22-34, 11_19_wallpushups_all_losses_test.yml:4,22-32
CONFIGS = {
    "1_17_ball_complex_L1_64": dict(tag="C1", dim_model=2048, num_heads=8, num_encoder_layers=4, num_decoder_layers=8,
                                    frame_size=64, dropout_p=0.1, frames_per_clip=5),
    "1_15_kitti_L1_64": dict(tag="C2", dim_model=2048, num_heads=8, num_encoder_layers=4, num_decoder_layers=8,
                             frame_size=64, dropout_p=0.1, frames_per_clip=5),
    "11_27_ucf_final": dict(tag="C3", dim_model=2048, num_heads=8, num_encoder_layers=4, num_decoder_layers=8,
                            frame_size=128, dropout_p=0.1, frames_per_clip=5),
    "11_20_wallpushups_dim_2048": dict(tag="C4", dim_model=2048, num_heads=8, num_encoder_layers=6,
                                       num_decoder_layers=6, frame_size=128, dropout_p=0.1, frames_per_clip=5),
    "11_19_wallpushups_all_losses_test": dict(tag="C5", dim_model=1024, num_heads=16, num_encoder_layers=12,
                                              num_decoder_layers=12, frame_size=128, dropout_p=0.1, frames_per_clip=5),
}


def latent_dim(frame_size, compression=8):
    """E = 4 * (F/8)^2, written as the reference computes it (models/transformer.py:37)."""
    return frame_size // compression * frame_size // compression * 4


def load_config(config_name):
    """utils/config.py:8-17: read ./config/<name>.yml relative to the cwd."""
    import yaml
    path = os.path.join("./config", config_name + ".yml")
    with open(path, "r") as stream:
        data = yaml.safe_load(stream)
    obj = SimpleNamespace(**data)
    obj.CONFIG_NAME = config_name
    return obj


def parse_config_args():
    """Same flags as utils/config.py:27-49 (--dataset and --config required)."""
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", type=str, required=True)
    p.add_argument("--save_best", type=bool, default=False)
    p.add_argument("--folder", type=str, default=None)
    p.add_argument("--config", type=str, required=True)
    p.add_argument("--resume", type=bool, default=False)
    p.add_argument("--debug", type=bool, default=False)
    p.add_argument("--flip", type=bool, default=False)
    p.add_argument("--pred_frames", type=int, default=1)
    p.add_argument("--show", type=bool, default=False)
    p.add_argument("--old_name", type=str, default="old_name_default")
    p.add_argument("--fullscreen", type=bool, default=False)
    p.add_argument("--save_output", type=bool, default=False)
    p.add_argument("--index", type=int, default=0)
    p.add_argument("--denoise", type=bool, default=False)
    p.add_argument("--mode", type=str, default="")
    p.add_argument("--denoise_start_step", type=int, default=40)
    args = p.parse_args()
    return load_config(args.config), args
