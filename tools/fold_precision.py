"""Precision of the folded LayerNorm (SDVG_LN_FOLD=1) against the plain path: C2 architecture, `mixed`, 1024 clips,
teacher-forced and free-running against the chunked CPU oracle on a 128-clip subset (two PE chunks), 3 frames;
max-rel per frame (the parity metric) and rms-rel (robust).  Run once per setting of SDVG_LN_FOLD."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import sdvg_b200
from oracle import rollout as R
from conftest import load_golden
from test_gpu_parity import ours_from

g = load_golden("c1_rollout")
m, ref = ours_from(g, "mixed")
n = 3
ctx = torch.randn(1024, 10, 256, generator=torch.Generator().manual_seed(99))
sub = torch.cat([torch.arange(0, 64), torch.arange(960, 1024)])
with torch.no_grad():
    want = R.chunked(lambda c: R.rollout_ref(ref, c, n, 5), ctx[sub])
    want64 = R.chunked(lambda c: R.rollout_ref(ref.double(), c.double(), n, 5), ctx[sub]).float()
teacher = torch.zeros(1024, n, 256)
teacher[sub] = want
tf = sdvg_b200.rollout(m, ctx.to("cuda"), n, 5, teacher=teacher.to("cuda")).cpu()[sub]
fr = sdvg_b200.rollout(m, ctx.to("cuda"), n, 5).cpu()[sub]
def rms(a, b):
    return [float(((a[:, i] - b[:, i]).pow(2).mean() / b[:, i].pow(2).mean()).sqrt()) for i in range(a.shape[1])]
print("SDVG_LN_FOLD =", os.environ.get("SDVG_LN_FOLD", "unset"))
print("  teacher-forced max-rel per frame", [f"{float(e):.2e}" for e in R.max_rel_per_frame(tf, want)], "rms-rel", [f"{e:.2e}" for e in rms(tf, want)])
print("  free-running   max-rel per frame", [f"{float(e):.2e}" for e in R.max_rel_per_frame(fr, want)], "rms-rel", [f"{e:.2e}" for e in rms(fr, want)])
print("  fp32 oracle vs float64 oracle (free-running) max-rel", [f"{float(e):.2e}" for e in R.max_rel_per_frame(want, want64)])
