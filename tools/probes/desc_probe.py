"""Does a tcgen05 shared-memory descriptor accept a start address that is not a multiple of the 1024-byte swizzle atom?
The 1-CTA GEMM reads its B operand (K-major, SWIZZLE_128B, one 128-byte row per B row) from row `shift` of the staged
tile, with the descriptor's base-offset field set to `base`; D[:, n] must then equal A . B[n + shift].
(Needed to serve the 9 taps of a 3x3 convolution as shifted views of ONE haloed activation tile.)
python tools/probes/desc_probe.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import torch
    from relax_vqa_b200 import ops
    shift = int(os.environ.get("B200VQA_GEMM_BSHIFT", "0"))
    ctx = ops.Context(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(128, 64, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(256, 64, device="cuda", generator=g) * 0.5).half()
    out = ops.gemm_f16(ctx, A, B, None, impl=0)
    ref = A.float() @ B.float().t()
    n = 256 - shift - 8
    err = (out[:, :n] - ref[:, shift:shift + n]).abs().max().item()
    print(f"shift {shift} base {os.environ.get('B200VQA_GEMM_BBASE', '0')}: max err over {n} columns = {err:.3e}", flush=True)
    ctx.close()
    sys.exit(0)
for shift in (0, 8, 1, 2, 3, 7, 9, 58):
    for base in sorted({0, shift & 7}):
        e = dict(os.environ); e["B200VQA_GEMM_BSHIFT"] = str(shift); e["B200VQA_GEMM_BBASE"] = str(base)
        r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-300:])
