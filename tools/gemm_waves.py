"""Fixed cost per launch vs cost per wave of the 2-CTA linear kernel: N = 256 (one column tile), K = 768,
M = 256 * 74 * waves.  T(waves) = fixed + waves * wave_time, per mode (full / no epilogue / barriers only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relax_vqa_b200 import ops

ctx = ops.Context(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 768
N = 256
for mode, name in (("0", "full"), ("1", "no-epilogue"), ("5", "tma-only"), ("3", "mma-only"), ("7", "barriers-only")):
    os.environ["B200VQA_GEMM_NOEPI"] = mode
    row = []
    for waves in (1, 2, 4, 8, 16):
        M = 256 * 74 * waves
        A = (torch.randn(M, K, device="cuda") * 0.5).half(); B = (torch.randn(N, K, device="cuda") * 0.5).half(); bias = torch.randn(N, device="cuda")
        for _ in range(3): ops.gemm_f16(ctx, A, B, bias, impl=2)
        torch.cuda.synchronize()
        # back-to-back launches (as in the pipeline): 20 launches between one pair of events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.gemm_f16(ctx, A, B, bias, impl=2)
        e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 20 * 1e3)
    per_wave = (row[-1] - row[-2]) / 8
    print(f"K={K} {name:14s} us per launch for 1,2,4,8,16 waves: " + " ".join(f"{t:7.1f}" for t in row) +
          f" | per wave {per_wave:.2f} us (peak {256*256*K*2*74/2.38e15*1e6:.2f}), fixed {row[0] - per_wave:.1f} us", flush=True)
