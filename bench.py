#!/usr/bin/env python
"""Benchmark of the ReLaX-VQA hot path: 1080p videos/s end to end (features + MLP).

    python bench.py --gpus N --steps K --warmup W            # this repo (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

    python bench.py --workload {540p-8s,1080p-10s,2160p-20s,lsvq-mix}   # the other BASELINE.json configs (default 1080p-10s)

A step = one pass of the hot path over one batch of `--clips` synthetic clips per GPU (default workload: LIVE-VQC-shaped,
1920x1080, 300 frames @29.97 -> 22 sampled pairs).  The global batch of world x clips videos is split by the
longest-processing-time sharder (relax_vqa_b200/sharding.py) and the per-video rows are gathered over NCCL in global
video order.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs -> (H, W, sampled pairs per clip, default clips per GPU per step, video_type for the score rescale,
#                           CPU-sample pairs of the reference arm).  SURVEY.md 8(d); pairs from frames / int(fps / 2).
WORKLOADS = {
    "540p-8s": dict(H=540, W=960, pairs=18, clips=16, video_type="konvid_1k", cpu_pairs=18),      # KoNViD-1k-shaped (configs[0], [1])
    "1080p-10s": dict(H=1080, W=1920, pairs=22, clips=4, video_type="live_vqc", cpu_pairs=11),    # LIVE-VQC-shaped (configs[2]) - the headline
    "2160p-20s": dict(H=2160, W=3840, pairs=43, clips=2, video_type="youtube_ugc", cpu_pairs=3),  # YouTube-UGC-2160P-shaped (configs[3])
    "lsvq-mix": dict(H=None, W=None, pairs=None, clips=48, video_type="lsvq_train", cpu_pairs=16),  # LSVQ mixed resolutions (configs[4])
}
# algorithmic figures per sampled pair
DENSE_GFLOP_PER_PAIR = 129.9            # 3 x ResNet-50 (8.178) + 3 x ViT-B/16 (35.126)
BW_BYTES_PER_PIXEL_PAIR = 335.0         # bandwidth stages (absdiff/patch sums, Farneback, colouring, resizes)


def metric_name(workload):
    return "videos_per_sec_" + workload.split("-")[0].replace("lsvq", "lsvq_mix")


def lsvq_specs(n, seed):
    """n (H, W, pairs) triples drawn (seeded, with the dataset's frequencies) from the LSVQ-train shape histogram
    (relax_vqa_b200/data/lsvq_train_shapes.csv <- metadata/LSVQ_TRAIN_metadata.csv, tools/make_lsvq_shapes.py)."""
    import numpy as np
    rows = np.loadtxt(os.path.join(ROOT, "relax_vqa_b200", "data", "lsvq_train_shapes.csv"), delimiter=",", skiprows=1, dtype=np.int64)
    rng = np.random.default_rng(seed)
    pick = rng.choice(len(rows), size=n, p=rows[:, 3] / rows[:, 3].sum())
    return [(int(rows[i, 1]), int(rows[i, 0]), int(rows[i, 2])) for i in pick]


def global_specs(workload, n):
    wl = WORKLOADS[workload]
    if workload == "lsvq-mix":
        return lsvq_specs(n, seed=2024)
    return [(wl["H"], wl["W"], wl["pairs"])] * n


def workload_config(workload, clips, world):
    wl = WORKLOADS[workload]
    cfg = dict(workload=workload, width=wl["W"], height=wl["H"], pairs_per_clip=wl["pairs"], clips_per_gpu_per_step=clips,
               parallelism=f"video-sharded x{world}")
    if workload == "lsvq-mix":
        specs = global_specs(workload, clips * world)
        cfg.update(width="mixed", height="mixed", pairs_per_clip="mixed", distinct_resolutions=len({(h, w) for h, w, _ in specs}),
                   mean_pairs=sum(p for _, _, p in specs) / len(specs), source="LSVQ_TRAIN shape histogram, seeded draw")
    return cfg


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (NVML, the source of nvidia-smi's numbers)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.marks = index, [], None, []

    def mark(self):
        """Called at the start and at the end of the timed region: only the samples in between are reported."""
        self.marks.append(len(self.rows))

    def start(self):
        """NVML from a polling thread of this process (what nvidia-smi itself reads, without a second process contending
        for the driver while kernels are being launched); nvidia-smi -lms as the fallback."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            idx = self.index
            vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
            if self.index < len(vis) and vis[self.index].isdigit():
                idx = int(vis[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            masks = (0x8, 0x40, 0x20, 0x4)              # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
            self.halt = threading.Event()

            def poll():
                while not self.halt.is_set():
                    try:
                        r = int(get_reasons(h))
                        self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)] +
                                         ["Active" if r & m else "Not Active" for m in masks])
                    except Exception:
                        pass
                    self.halt.wait(0.02)

            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            self.proc = "nvml"
            return
        except Exception:
            self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)                                 # let the sample that covers the end of the region arrive (100 ms polling:
                                                         # faster polling makes nvidia-smi contend with the kernel launches)
        if self.proc == "nvml":
            self.halt.set()
        else:
            self.proc.terminate()
        self.t.join(timeout=2)
        if len(self.marks) == 2:
            lo, hi = self.marks[0], max(self.marks[1] + 1, self.marks[0] + 1)
            self.rows = self.rows[lo:hi] or self.rows[-1:]
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                    source="nvml" if self.proc == "nvml" else "nvidia-smi")



def cpu_reference_sample(workload, n_pairs, threads=None):
    """Times the CPU restatement of the reference (oracle/) on `n_pairs` sampled pairs + their full frames of `workload`.
    Uses cv2's Farneback (the reference's own dependency) for the flow stage.  -> (seconds per pair, pairs per clip)."""
    import cv2
    import torch
    from oracle import pipeline as P
    from relax_vqa_b200 import synth, weights
    if threads:
        torch.set_num_threads(threads)
    rsd, vsd = weights.seeded_resnet50_state_dict(), weights.seeded_vitb16_state_dict()
    if workload == "lsvq-mix":
        specs = global_specs(workload, 4096)
        mean_pairs = sum(p for _, _, p in specs) / len(specs)
        sample = specs[:n_pairs]                          # one pair from each of n_pairs drawn videos: time ~ the pair-weighted mix
        clips = [synth.make_clip(i, h, w, 1) for i, (h, w, _) in enumerate(sample)]
    else:
        wl = WORKLOADS[workload]
        mean_pairs = wl["pairs"]
        clips = [synth.make_clip(0, wl["H"], wl["W"], n_pairs)]
    flow = lambda a, b: cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    t0 = time.perf_counter()
    for fr, nx in clips:
        blocks = P.video_feature_blocks(fr, nx, rsd, vsd, flow_fn=flow)
        P.video_vector(blocks)
    return (time.perf_counter() - t0) / n_pairs, mean_pairs


def cpu_sample_pairs(workload):
    return int(os.environ.get("B200VQA_CPU_SAMPLE_PAIRS", WORKLOADS[workload]["cpu_pairs"]))   # ~10-30 s of host work per sample on 16 cores


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (Python reference -> the oracle port,
    one backbone forward per image instead of the reference's 15, i.e. favourable to the CPU), all host threads.
    A step = a bounded sample of sampled pairs of the workload (half a 1080p clip by default); K steps, capped at ~150 s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = args.workload
    n = cpu_sample_pairs(wl)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(wl, 1)
    times, t_start, mean_pairs = [], time.perf_counter(), None
    for _ in range(max(1, args.steps)):
        per_pair, mean_pairs = cpu_reference_sample(wl, n)
        times.append(per_pair)
        if time.perf_counter() - t_start > 150.0:
            break
    per_pair = sum(times) / len(times)
    vps = 1.0 / (per_pair * mean_pairs)
    clips = args.clips or WORKLOADS[wl]["clips"]
    line = dict(impl="reference", metric=metric_name(wl), value=vps, unit="videos/s", n_gpus=args.gpus, steps=len(times),
                warmup=min(args.warmup, 1), ms_per_step=per_pair * n * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", config=workload_config(wl, clips, args.gpus),
                cpu_baseline=dict(value=vps, unit="videos/s", cores=cores, kind="port",
                                  sample=f"{n} sampled pairs (+ their full frames) of the {wl} workload per step through oracle/pipeline.py with "
                                         f"cv2 Farneback; per-clip time = {mean_pairs:.2f} x per-pair time"),
                e2e=dict(value=vps, unit="videos/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="1080p-10s", choices=sorted(WORKLOADS))
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU per step (0 = the workload's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gemm-impl", type=int, default=0, help="0 default (2-CTA linears), 2 = 1-CTA kernel everywhere (A/B)")
    ap.add_argument("--gemm-sms", type=int, default=-1, help="SMs of the persistent tcgen05 grids (-1 = engine default, 0 = all)")
    ap.add_argument("--no-pipeline", action="store_true", help="A/B: join the image stages of a step before its backbones start AND "
                                                               "the previous step before the next one's image stages")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from relax_vqa_b200 import sharding, weights
    from relax_vqa_b200.engine import Engine, bind_host_to_gpu, synthetic_clips_on_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpus = bind_host_to_gpu(local)                      # pinned staging buffers NUMA-local to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]
    n_clips = args.clips or wl["clips"]
    video_type = wl["video_type"]
    eng = Engine(local, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
    if args.gemm_impl:
        eng.ctx.set_gemm_impl(args.gemm_impl)
    if args.gemm_sms >= 0:
        eng.set_gemm_sms(args.gemm_sms)
    if args.no_pipeline:
        eng.pipeline = False
    # ---- the global batch (world x clips videos) and this rank's share of it: the real sharder, also when all costs are equal
    specs = global_specs(args.workload, n_clips * world)
    costs = [sharding.video_cost(p, h, w) for h, w, p in specs]
    plan = sharding.shard_videos(costs, world)
    mine = plan[rank]
    clips = []
    for gi in mine:
        h, w, p = specs[gi]
        clips += synthetic_clips_on_device(1, h, w, p, eng.device, seed=1000 + gi)
    my_pairs = sum(specs[gi][2] for gi in mine)
    stream = torch.cuda.current_stream()

    def gather(feats, score):
        if world == 1:
            return feats, score
        return sharding.gather_rows(feats, plan), sharding.gather_rows(score.reshape(-1, 1), plan).reshape(-1)

    def step():
        feats, score = eng.predict(clips, video_type)
        return gather(feats, score)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                  # nvidia-smi needs a few 100 ms to come up: start it before the warm-up
    for _ in range(warmup):
        step()
    barrier()
    sampler.mark()
    l0 = eng.launches
    e0, e1, e_local = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        feats_l, score_l = eng.predict(clips, video_type)
        if i + 1 == args.steps:
            e_local.record(stream)                       # this rank's own finish time (before the last gather): the load balance of the plan
        feats, score = gather(feats_l, score_l)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=eng.device)
    local_ms = torch.tensor([e0.elapsed_time(e_local)], device=eng.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        parts = [torch.empty_like(local_ms) for _ in range(world)]
        dist.all_gather(parts, local_ms)
        rank_ms = [float(p.item()) for p in parts]
    else:
        rank_ms = [float(local_ms.item())]
    ms = float(ms.item())
    launches = eng.launches - l0
    sampler.mark()
    clocks = sampler.stop() if rank == 0 else None
    videos = world * n_clips * args.steps
    value = videos / (ms / 1e3)

    # ---- self check: a video scored inside the batch == the same video scored alone, bit for bit (batch-invariant kernels),
    # and the gathered matrix holds this rank's rows at their global positions
    f1, s1 = eng.predict([clips[0]], video_type)
    torch.cuda.synchronize()
    self_check = bool(torch.equal(s1[0], score_l[0]) and torch.equal(f1[0], feats_l[0]) and torch.equal(feats[mine[0]], feats_l[0])
                      and bool(torch.isfinite(score).all()) and feats.shape == (world * n_clips, 35203))
    del f1, s1

    # ---- e2e: same metric through the public API with HOST (pinned) buffers, H2D + D2H inside the timed region
    host = [(c.frames.cpu().pin_memory(), c.nexts.cpu().pin_memory()) for c in clips]
    h2d = sum(f.numel() + n.numel() for f, n in host)
    def e2e_loop(n):
        ticket = eng.submit_host(host, video_type)      # two batches in flight: the copies of step k+1 run under step k's kernels
        for i in range(n):
            nxt = eng.submit_host(host, video_type) if i + 1 < n else None
            out = eng.result(ticket)                     # scores of step i on the host
            ticket = nxt
        return out

    e2e_loop(max(warmup, 3))                             # same pattern as the timed loop (device staging blocks of both in-flight batches exist)
    barrier()
    e2e_steps = max(2, args.steps)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    f_, s_ = e2e_loop(e2e_steps)
    t1.record(stream)
    barrier()
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], device=eng.device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * n_clips * e2e_steps / (float(e2e_ms.item()) / 1e3)
    self_check = self_check and bool(torch.equal(s_.to(score_l.device), score_l))        # host-buffer entry == device-buffer entry

    # ---- roofline of the dominant kernel (tcgen05 GEMM / implicit-GEMM conv): CUDA events around every launch
    peaks = measured_peaks()
    eng.concurrent = False            # per-launch events need the kernels one after the other (in the timed steps the
    eng.set_profiling(True)           # clips' image stages and the two backbones share the GPU on several streams)
    eng.ctx.profile_read()
    eng.profile_read_flow()
    for _ in range(2):
        eng.predict(clips, video_type)
    gemm_ms, gemm_launches, gemm_flops = eng.ctx.profile_read()
    flow_ms, flow_launches, flow_bytes = eng.profile_read_flow()
    # the same per-launch timing split by kernel: the ViT linears (SM-pair kernel) and the ResNet convolutions (implicit GEMM,
    # HBM-bound on the 56x56 / 28x28 layers) on a batch of one step's backbone images
    from relax_vqa_b200 import ops as _ops
    n_img = max(1, min(3 * my_pairs, 512))
    probe = torch.randint(0, 256, (n_img, 224, 224, 3), dtype=torch.uint8, device=eng.device)
    by_kernel = {}
    for name, fn in (("gemm_tcgen05_kernel (ResNet-50 convolutions)", lambda: _ops.resnet50_features(eng.ctx, probe, is_bgr=True, want_stack=True, want_pool=True)),
                     ("gemm2cta_tcgen05_kernel (ViT-B/16 linears)", lambda: _ops.vitb16_features(eng.ctx, probe, is_bgr=True))):
        fn()
        eng.ctx.profile_read()
        fn()
        k_ms, k_n, k_fl = eng.ctx.profile_read()
        k_tf = k_fl / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
        by_kernel[name] = dict(achieved=k_tf, frac=k_tf / peaks["tf_sustained"], launches=k_n, ms=k_ms, images=n_img)
    del probe
    eng.set_profiling(False)
    eng.concurrent = True
    step_ms = ms / args.steps
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp) and args.workload == "1080p-10s":          # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic = json.load(open(tp))["traffic_bytes_per_launch"]
    roofline = dict(bound="tensor", kernel="gemm_tcgen05_kernel", achieved=achieved, peak=peaks["tf_sustained"], unit="TFLOP/s",
                    frac=achieved / peaks["tf_sustained"], traffic=traffic, peak_source=peaks["src"] + " (sustained fp16/bf16 dense)",
                    launches_per_step=gemm_launches // 2, kernel_ms_per_step=gemm_ms / 2, share_of_step=(gemm_ms / 2) / step_ms,
                    algorithmic_gflop_per_pair=gemm_flops / 2 / max(my_pairs, 1) / 1e9, by_kernel=by_kernel)
    flow_gbs = flow_bytes / (flow_ms / 1e3) / 1e9 if flow_ms > 0 else 0.0
    ratio, fsrc = 1.0, None
    fp = os.path.join(ROOT, "profiles", "flow_traffic.json")
    if os.path.exists(fp):          # dram bytes / algorithmic bytes of the level-0 launch from the committed ncu capture
        fj = json.load(open(fp))
        ratio, fsrc = fj["dram_over_algorithmic"], "profiles/flow_traffic.json"
    roofline_hbm = dict(bound="hbm", kernel=eng.flow_kernel_name(), achieved=flow_gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=flow_gbs / peaks["hbm_gbs"],
                        traffic=(flow_bytes / max(flow_launches, 1)) * ratio if fsrc else None, launches_per_step=flow_launches // 2,
                        kernel_ms_per_step=flow_ms / 2, share_of_step=(flow_ms / 2) / step_ms,
                        note="largest single bandwidth kernel (Farneback iteration); algorithmic bytes per launch as counted by the library "
                             "(DESIGN.md 4.2); traffic = that x the ncu dram/algorithmic ratio of the level-0 launch (%s)" % fsrc)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n = cpu_sample_pairs(args.workload)
        per_pair, mean_pairs = cpu_reference_sample(args.workload, n)
        cpu = dict(value=1.0 / (per_pair * mean_pairs), unit="videos/s", cores=torch.get_num_threads(), kind="port",
                   sample=f"{n} sampled pairs + their full frames of the {args.workload} workload through oracle/pipeline.py (cv2 Farneback), "
                          f"{per_pair * n:.1f} s; clip = {mean_pairs:.2f} pairs")
    cfg = workload_config(args.workload, n_clips, world)
    line = dict(metric=metric_name(args.workload), value=value, unit="videos/s", n_gpus=world, steps=args.steps, warmup=warmup,
                ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16", data="synthetic",
                config=cfg, value_is="device-resident: frames already in HBM (e2e.value is the host-to-score number)",
                timing="inputs larger than L2 (%.0f MB per step per GPU); CUDA events, max over ranks" % (h2d / 1e6),
                clocks=clocks, gpu_launches=int(launches), self_check=self_check,
                e2e=dict(value=e2e_value, unit="videos/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(4 * len(clips)),
                         host_cpus_bound=len(cpus) if cpus else None),
                rank_ms_per_step=[t / args.steps for t in rank_ms], sharder="LPT (sharding.shard_videos) + ragged all-gather (sharding.gather_rows)",
                engine=dict(pipeline=bool(eng.pipeline), gemm_sms=eng.gemm_sms, lanes=eng.LANES),
                roofline=roofline, roofline_hbm=roofline_hbm, cpu_baseline=cpu)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not self_check:
        sys.exit("bench self-check failed: batched and single-clip results differ")


if __name__ == "__main__":
    main()
