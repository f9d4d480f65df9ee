"""C1 (BASELINE configs[0]: d2048 4e/8d, 8 clips, window 5) through the per-kernel launch chain - the workload behind the
`c1_w5` extra of bench.py - for an ncu launch list:
    SDVG_PK=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c1_launches.csv python tools/c1_chain.py [precision] [graph 0|1]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdvg_b200

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
if len(sys.argv) > 2:
    os.environ["SDVG_GRAPH"] = sys.argv[2]
os.environ.setdefault("SDVG_PK", "0")
cfg = sdvg_b200.CONFIGS["1_17_ball_complex_L1_64"]
B, C, P, W = int(os.environ.get("C1_B", "8")), 10, 10, int(os.environ.get("C1_W", "5"))
torch.manual_seed(0)
m = sdvg_b200.Transformer(0, cfg["dim_model"], cfg["num_heads"], cfg["num_encoder_layers"], cfg["num_decoder_layers"], 0.1,
                          frame_size=64, precision=prec, max_clips=B, max_tokens=int(os.environ.get("C1_MAXTOK", "10")), max_history=C + P).eval().cuda()
ctx = torch.randn(B, C, 256, generator=torch.Generator().manual_seed(1234)).cuda()
pe = sdvg_b200.pe_index_for(0, B, ctx.device) if os.environ.get("C1_PE") else None
out = m.rollout(ctx, P, W, pe_index=pe)
for _ in range(3):
    m.rollout(ctx, P, W, out=out, pe_index=pe)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m.rollout(ctx, P, W, out=out, pe_index=pe)
e1.record()
torch.cuda.synchronize()
print(f"C1 {prec} B={B} SDVG_PK={os.environ.get('SDVG_PK')}: {e0.elapsed_time(e1) / 5 / P * 1e3:.1f} us per pass")
