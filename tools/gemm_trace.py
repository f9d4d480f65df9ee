"""Pipeline timeline of CTA 0 of the CTA-pair GEMM (%globaltimer stamps):  python tools/gemm_trace.py M N K prec bn"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdvg_b200

M, N, K = (int(v) for v in sys.argv[1:4])
prec = sys.argv[4] if len(sys.argv) > 4 else "fp16"
bn = int(sys.argv[5]) if len(sys.argv) > 5 else -192
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g)
W = torch.randn(N, K, device="cuda", generator=g) * 0.05
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
heat = int(sys.argv[6]) if len(sys.argv) > 6 else 3
sdvg_b200.gemm(A, W, None, precision=prec, block_n=bn, iters=heat)
os.environ["SDVG_TRACE_BUF"] = str(buf.data_ptr())
_, ms = sdvg_b200.gemm(A, W, None, precision=prec, block_n=bn, iters=1)
torch.cuda.synchronize()
t = buf.cpu().tolist()
t0 = t[0]
names = {0: "entry", 1: "prologue done", 2: "pdl_wait passed", 3: "first TMA issued", 4: "main work done (CTA0)", 5: "exit"}
print(f"{prec} {M}x{N}x{K} bn={bn}: event-timed {ms*1e3:.1f} us (single launch incl. launch latency)")
for i in (0, 1, 2, 3):
    if t[i]:
        print(f"  {names[i]:28s} {(t[i]-t0)/1e3:8.2f} us")
for i in range(8):
    if t[8 + i] and t[16 + i]:
        print(f"  tile {i}: first stage full {(t[8+i]-t0)/1e3:8.2f}  mma issued {(t[16+i]-t0)/1e3:8.2f}  "
              f"acc ready {(t[24+i]-t0)/1e3:8.2f}  epilogue done {(t[32+i]-t0)/1e3:8.2f} us")
for i in (4, 5):
    print(f"  {names[i]:28s} {(t[i]-t0)/1e3:8.2f} us")
print(f"  SM clock during the kernel: {(t[41]-t[40])/(t[5]-t[0])*1e3:.0f} MHz (after {heat} warm-up launches)")
