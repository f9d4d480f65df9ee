"""CPU: the JSON-line contract of bench.py / bench_train.py, exercised through their reference arms (the oracle port on
the host cores - the only arm that runs without a GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def run(args):
    out = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=900, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def check(d, metric, kinds=("port",)):
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == metric and d["value"] > 0 and d["higher_is_better"] is True
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in kinds and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]


def test_rollout_reference_arm_line():
    d = run(["bench.py", "--impl", "reference", "--config", "1_17_ball_complex_L1_64", "--steps", "1", "--warmup", "1",
             "--cpu-sample-steps", "1"])
    # "reference" when baseline/_ref (the unmodified reference module, oracle/install_ref.py) is installed, else the port
    check(d, "latent_frames_per_sec_rollout", kinds=("reference", "port"))
    from oracle import install_ref
    assert d["cpu_baseline"]["kind"] == ("reference" if install_ref.available() else "port")
    assert d["unit"] == "latent frames/s"


def test_training_reference_arm_line():
    d = run(["bench.py", "--workload", "train", "--impl", "reference", "--steps", "1", "--warmup", "1"])
    check(d, "train_clips_per_sec")
    assert "dropout" in d["config"]["workload"]
