"""Data-parallel training check (run under torchrun on >= 2 GPUs): each rank trains on its shard of a global batch;
after the NCCL all-reduce the averaged gradients must equal the single-process gradients of the global batch
(oracle on CPU, rank 0), replicas must stay bit-identical, and the overlapped (per-bucket, announced by the library
while the backward pass is enqueued) and plain all-reduce must agree."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
import sdvg_b200
from oracle import train as OT
from oracle.ref_module import RefTransformer

KW = dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07, lambda_contrastive=0.1)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per = 4
    B = per * world
    batches = [OT.make_batch(B, 6, 256, seed=40 + s) for s in range(2)]
    results = {}
    for overlap in (True, False):
        torch.manual_seed(3)
        m = sdvg_b200.Transformer(0, 128, 4, 2, 2, 0.0, frame_size=64, precision="fp32").to(dev)
        tr = sdvg_b200.AdamTrainer(m, lr=1e-3, frames_to_predict=5, overlap=overlap, **KW)
        pe = torch.arange(rank * per, (rank + 1) * per)
        for s in range(2):
            tr.step(batch := batches[s][rank * per:(rank + 1) * per].to(dev), pe_index=pe)
            if s == 0:
                flat, split = tr.gradients(dev)
                g0 = {k: (tr.gradient(k) / world).cpu() for k, _ in m.named_parameters()}
        tr.pull_weights()
        results[overlap] = (g0, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()})
        # replicas identical?
        chk = torch.stack([v.double().sum() for v in m.state_dict().values()]).to(dev)
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(torch.equal(allc[0], c) for c in allc), "replicas diverged"
    ga, wa = results[True]; gb, wb = results[False]
    # the overlapped path reduces finer ranges than the plain one, and NCCL's summation order depends on the position
    # of an element inside the reduced buffer: equal up to fp32 re-association, not bit for bit
    for k in ga:
        assert float((ga[k] - gb[k]).abs().max()) <= 1e-6 * float(gb[k].abs().max()) + 1e-12, ("overlap changes gradients", k)
    if rank == 0:
        torch.manual_seed(3)
        ref = RefTransformer(0, 128, 4, 2, 2, 0.0, frame_size=64)
        opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
        _, _, grads = OT.train_step_ref(ref, opt, batches[0], 5, **KW)
        worst = max(float((ga[k] - g).abs().max() / g.abs().max()) for k, g in grads.items())
        print(f"DP check: world {world}, averaged gradients vs global-batch oracle: worst rel err {worst:.2e}")
        assert worst <= 1e-4, worst
        print("DP check OK (replicas identical, overlapped == plain all-reduce up to fp32 re-association)")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
