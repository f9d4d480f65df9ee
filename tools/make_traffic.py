"""profiles/r2_traffic.json + a compact per-launch summary from an `ncu --set full` raw CSV of the bench's GEMM kernels.

    ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel --launch-skip 700 -c 6 \
        -o gpurun_out/r2_gemm_full python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline
    ncu -i gpurun_out/r2_gemm_full.ncu-rep --page raw --csv > gpurun_out/r2_gemm_ncu_full_raw.csv
    python tools/make_traffic.py gpurun_out/r2_gemm_ncu_full_raw.csv [git_head_of_the_capture]

bench.py (read_traffic) drops the number when the kernel sources hashed here changed or the commit is not an ancestor of HEAD.
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["sd-video-gen_b200/csrc/gemm_tc.cuh", "sd-video-gen_b200/csrc/gemm_tc2.cuh", "sd-video-gen_b200/csrc/ptx.cuh",
           "sd-video-gen_b200/csrc/common.cuh"]
COLS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def to_bytes(v, unit):
    return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    raw = sys.argv[1]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {c: hdr.index(c) for c in COLS if c in hdr}
    out = os.path.join(ROOT, "profiles", os.path.basename(raw).replace("_raw", ""))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(list(idx))
        w.writerow([units[i] for i in idx.values()])
        for r in data:
            w.writerow([r[i] for i in idx.values()])
    print("wrote", out)
    # dominant instantiation of the C2 step: gemm_tc2_kernel<192,0,0> at M=5120 N=2048 K=2048 (the launches that read ~71 MB)
    big = [r for r in data if "gemm_tc2_kernel<192, 0, 0>" in r[idx["Kernel Name"]]
           and to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) > 50e6]
    if not big:
        return
    rd = sum(to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) for r in big) / len(big)
    wr = sum(to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in big) / len(big)
    h = hashlib.sha256()
    for s in SOURCES:
        h.update(open(os.path.join(ROOT, s), "rb").read())
    head = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    M, N, K = 5120, 2048, 2048
    tj = {
        "source": f"profiles/{os.path.basename(out)} (ncu --set full --clock-control none, cold L2, B200; mean of {len(big)} launches)",
        "kernel": "sdvg::gemm_tc2_kernel<192,false,false> M=5120 N=2048 K=2048 (52 of 74 GEMM launches per model pass)",
        "git_head": head, "kernel_sources": SOURCES, "kernel_sources_sha256": h.hexdigest(),
        "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
        # A planes 2 B + W planes 2 B + fp32 residual in + fp32 out + 16-bit planes of the pre-norm sums (out-proj / FF2
        # with the LayerNorm folded in; DESIGN.md kernel table)
        "algorithmic_bytes_per_launch": 2 * M * K + 2 * N * K + 4 * M * N + 4 * M * N + 2 * M * N,
        "tensor_pipe_active_pct": sum(float(r[idx["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]]) for r in big) / len(big),
        "duration_us": sum(float(r[idx["gpu__time_duration.sum"]]) for r in big) / len(big),
    }
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
        json.dump(tj, f, indent=1)
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
