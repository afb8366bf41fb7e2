"""A/B of the 2-CTA (impl 2) and 4-CTA multicast (impl 3) linear kernels on the ViT shapes of a 264-image pass
(CUDA events, L2 flushed between runs).  python tools/gemm_mc_ab.py [images]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relax_vqa_b200 import ops

images = int(sys.argv[1]) if len(sys.argv) > 1 else 264
M = images * 197
ctx = ops.Context(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for name, N, K in [("qkv", 2304, 768), ("proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072), ("8192^3", 8192, 8192)]:
    m = 8192 if name == "8192^3" else M
    A = (torch.randn(m, K, device="cuda") * 0.5).half(); B = (torch.randn(N, K, device="cuda") * 0.5).half(); bias = torch.randn(N, device="cuda")
    line = f"{name:7s} M={m} N={N} K={K}:"
    outs = {}
    for impl in (2, 3):
        for _ in range(3):
            outs[impl] = ops.gemm_f16(ctx, A, B, bias, impl=impl)
        ts = []
        for _ in range(9):
            flush.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gemm_f16(ctx, A, B, bias, impl=impl); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort(); t = ts[len(ts) // 2]
        line += f"  impl {impl}: {t * 1e3:7.1f} us {2.0 * m * N * K / (t * 1e-3) / 1e12:7.1f} TFLOP/s"
    line += f"  identical={bool(torch.equal(outs[2], outs[3]))}"
    print(line, flush=True)
ctx.close()
