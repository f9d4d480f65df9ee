"""SDVG_PK_TRACE=n,cta,first python tools/pk_trace_rollout.py [precision] [window]: op-level pipeline trace of a C1 rollout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdvg_b200
prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = sdvg_b200.CONFIGS["1_17_ball_complex_L1_64"]
B, C, P = int(os.environ.get("C1_B", "8")), 10, 10
torch.manual_seed(0)
m = sdvg_b200.Transformer(0, cfg["dim_model"], cfg["num_heads"], cfg["num_encoder_layers"], cfg["num_decoder_layers"], 0.1,
                          frame_size=64, precision=prec, max_clips=B, max_tokens=10, max_history=C + P).eval().cuda()
ctx = torch.randn(B, C, 256, generator=torch.Generator().manual_seed(1234)).cuda()
out = m.rollout(ctx, P, W)
for _ in range(int(os.environ.get('PK_REPEAT', '0'))):
    m.rollout(ctx, P, W, out=out)
torch.cuda.synchronize()
