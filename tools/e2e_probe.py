"""Where does the end-to-end (host buffers) time go?  Pure H2D rate, compute alone, sync and pipelined loops."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relax_vqa_b200 import weights
from relax_vqa_b200.engine import Engine, bind_host_to_gpu, synthetic_clips_on_device
if os.environ.get("BIND"):
    print("bound to", len(bind_host_to_gpu(0) or []), "cpus of", os.cpu_count(), flush=True)

H, W, PAIRS, CLIPS, STEPS = 1080, 1920, 22, 4, 8
eng = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
clips = synthetic_clips_on_device(CLIPS, H, W, PAIRS, eng.device, seed=1)
host = [(c.frames.cpu().pin_memory(), c.nexts.cpu().pin_memory()) for c in clips]
nbytes = sum(f.numel() + n.numel() for f, n in host)
main = torch.cuda.current_stream()

def timed(fn, n=STEPS):
    fn(3); torch.cuda.synchronize()
    t = time.perf_counter(); fn(n); torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3

def copies(n):
    for _ in range(n):
        d = [(f.to("cuda", non_blocking=True), x.to("cuda", non_blocking=True)) for f, x in host]
def compute(n):
    for _ in range(n): eng.predict(clips, "live_vqc")
def sync_loop(n):
    for _ in range(n): eng.predict_host(host, "live_vqc")
def pipe_loop(n):
    t = eng.submit_host(host, "live_vqc")
    for i in range(n):
        nx = eng.submit_host(host, "live_vqc") if i + 1 < n else None
        eng.result(t); t = nx
ms = timed(copies); print(f"H2D only: {ms:.1f} ms/step = {nbytes / ms / 1e6:.1f} GB/s", flush=True)
print(f"compute only: {timed(compute):.1f} ms/step", flush=True)
print(f"sync loop: {timed(sync_loop):.1f} ms/step", flush=True)
print(f"pipelined loop: {timed(pipe_loop):.1f} ms/step", flush=True)
print("allocator:", torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], "device allocs;", torch.cuda.memory_reserved() / 1e9, "GB reserved")
print(f"pipelined loop again: {timed(pipe_loop):.1f} ms/step", flush=True)
print("allocator:", torch.cuda.memory_stats()["num_device_alloc"], "device allocs;", torch.cuda.memory_reserved() / 1e9, "GB reserved")

# ---- (a) compute on resident clips while unrelated H2D copies run on a side stream: contention only
side = torch.cuda.Stream()
def compute_with_background_copies(n):
    for _ in range(n):
        with torch.cuda.stream(side):
            d = [(f.to("cuda", non_blocking=True), x.to("cuda", non_blocking=True)) for f, x in host]
        eng.predict(clips, "live_vqc")
print(f"compute + unrelated background copies: {timed(compute_with_background_copies):.1f} ms/step", flush=True)
# ---- (b) device time of the compute part inside the pipelined loop
evs = []
_orig = eng.predict
def predict_timed(c, vt=None, is_finetune=False):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = _orig(c, vt, is_finetune); b.record(); evs.append((a, b)); return out
eng.predict = predict_timed
t = timed(pipe_loop)
torch.cuda.synchronize()
inner = [a.elapsed_time(b) for a, b in evs[-STEPS:]]
gaps = [evs[i][1].elapsed_time(evs[i + 1][0]) for i in range(len(evs) - STEPS, len(evs) - 1)]
print(f"pipelined: {t:.1f} ms/step; compute part {sum(inner) / len(inner):.1f} ms; gap between steps {sum(gaps) / len(gaps):.2f} ms", flush=True)

# ---- (c) compute on clips that were copied from the host once and then stay resident
eng.predict = _orig
from relax_vqa_b200.engine import Clip
copied = [Clip(f.to("cuda"), x.to("cuda")) for f, x in host]
torch.cuda.synchronize()
def compute_copied(n):
    for _ in range(n): eng.predict(copied, "live_vqc")
print(f"compute on once-copied resident clips: {timed(compute_copied):.1f} ms/step; equal to generated: "
      f"{all(torch.equal(a.frames, b.frames) and torch.equal(a.nexts, b.nexts) for a, b in zip(copied, clips))}", flush=True)
print(f"compute on generated clips again: {timed(compute):.1f} ms/step", flush=True)
for c in (clips[0], copied[0]):
    print("ptrs", hex(c.frames.data_ptr()), hex(c.nexts.data_ptr()), c.frames.is_contiguous(), c.frames.stride())

# ---- (d) host-side time of submit / result in the pipelined loop, and copy-done vs compute-start on the device
import statistics
sub, res = [], []
t = eng.submit_host(host, "live_vqc")
for i in range(STEPS):
    t0 = time.perf_counter(); nx = eng.submit_host(host, "live_vqc"); t1 = time.perf_counter()
    eng.result(t); t2 = time.perf_counter(); t = nx
    sub.append((t1 - t0) * 1e3); res.append((t2 - t1) * 1e3)
eng.result(t)
print("host ms per submit:", [round(v, 1) for v in sub], "per result wait:", [round(v, 1) for v in res], flush=True)
