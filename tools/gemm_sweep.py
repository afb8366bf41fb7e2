"""Timing experiments on the tcgen05 GEMM kernels (CUDA events, L2 flushed between runs)."""
import os, sys, itertools, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relax_vqa_b200 import ops

def run(M, N, K, impl, reps=10):
    ctx = ops.Context(0)
    A = (torch.randn(M, K, device="cuda") * 0.5).half(); B = (torch.randn(N, K, device="cuda") * 0.5).half(); bias = torch.randn(N, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3): ops.gemm_f16(ctx, A, B, bias, impl=impl)
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_f16(ctx, A, B, bias, impl=impl); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2]
    return t * 1e3, 2.0 * M * N * K / (t * 1e-3) / 1e12

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        M, N, K, impl = map(int, sys.argv[2:6])
        us, tf = run(M, N, K, impl)
        print(json.dumps(dict(M=M, N=N, K=K, impl=impl, us=round(us, 1), tflops=round(tf, 1), env={k: v for k, v in os.environ.items() if k.startswith("B200VQA_GEMM")})))
    else:
        shapes = [(25088, 2304, 768), (8192, 8192, 8192)]
        for (M, N, K) in shapes:
            # NOEPI bits: 1 = skip epilogue, 2 = skip TMA loads, 4 = skip MMA issue (pipeline-skeleton experiments)
            for impl, env in [(0, {}), (0, {"B200VQA_GEMM_NOEPI": "1"}), (0, {"B200VQA_GEMM_NOEPI": "3"}), (0, {"B200VQA_GEMM_NOEPI": "5"}),
                              (0, {"B200VQA_GEMM_NOEPI": "7"}), (2, {}), (2, {"B200VQA_GEMM_NOEPI": "1"}), (2, {"B200VQA_GEMM_NOEPI": "3"}),
                              (2, {"B200VQA_GEMM_NOEPI": "5"}), (2, {"B200VQA_GEMM_NOEPI": "7"})]:
                e = dict(os.environ); e.update(env)
                out = subprocess.run([sys.executable, __file__, "one", str(M), str(N), str(K), str(impl)], env=e, capture_output=True, text=True)
                print(out.stdout.strip() or out.stderr[-300:])
