"""Batched fast path: sampled frame pairs -> 35,203-dim video vectors -> MOS, all on one B200.

This is the call a user makes (``Engine.extract`` / ``Engine.predict``); the reference-named
per-image functions in ``main_fragment_layerstack`` etc. are thin views over the same kernels.
Stage order follows src/demo_test.py:76-219; every stage is a C-ABI call (include/b200vqa.h).
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops, weights


def bind_host_to_gpu(device_index=0):
    """Pins the calling process to the CPUs next to GPU `device_index` (NVML affinity mask), so that pinned host buffers
    allocated afterwards are NUMA-local to that GPU: remote pinned memory halves the H2D rate (about 25 instead of
    55 GB/s on the 2-socket B200 hosts).  Returns the CPU set, or None if NVML / the scheduler call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


@dataclass
class Clip:
    """Sampled frames of one video, BGR uint8 (what cv2.imread returns for the reference's PNGs).

    frames: [Tf, H, W, 3]  frames n with n % k == 0        (process_video,  vf_extract.py:103-111)
    nexts : [Tp, H, W, 3]  frames n with (n-1) % k == 0    (process_video_residual, :113-121)
    Pair i is (frames[i], nexts[i]); Tp <= Tf."""
    frames: torch.Tensor
    nexts: torch.Tensor
    ready: Optional[torch.cuda.Event] = None     # recorded on the copy stream when the frames were staged asynchronously


class Engine:
    def __init__(self, device=0, resnet_sd=None, vit_sd=None, head_sd=None, imputer_mean=None, scaler_scale=None,
                 scaler_min=None, seed_if_missing=False):
        """seed_if_missing=True (bench / tests / smoke only) substitutes seeded synthetic backbones for missing state
        dicts; by default a missing backbone is an error, as features from random weights are meaningless."""
        if (resnet_sd is None or vit_sd is None) and not seed_if_missing:
            raise ops._lib.B200VQAError("weights not loaded (B200VQA_ENOTLOADED): Engine needs resnet_sd and vit_sd "
                                        "(seed_if_missing=True substitutes seeded synthetic weights)")
        self.ctx = ops.Context(device)
        self.device = self.ctx.device
        self.synthetic_weights = resnet_sd is None or vit_sd is None
        if resnet_sd is None:
            resnet_sd = weights.seeded_resnet50_state_dict()
        if vit_sd is None:
            vit_sd = weights.seeded_vitb16_state_dict()
        ops.load_resnet50(self.ctx, resnet_sd)
        ops.load_vitb16(self.ctx, vit_sd)
        self.has_head = False
        if head_sd is not None:
            self.load_head(head_sd, imputer_mean, scaler_scale, scaler_min)

    def load_head(self, head_sd, imputer_mean=None, scaler_scale=None, scaler_min=None):
        n = weights.FEATURE_DIM
        imputer_mean = np.zeros(n) if imputer_mean is None else imputer_mean
        scaler_scale = np.ones(n) if scaler_scale is None else scaler_scale
        scaler_min = np.zeros(n) if scaler_min is None else scaler_min
        ops.load_head(self.ctx, head_sd, imputer_mean, scaler_scale, scaler_min)
        self.has_head = True

    def close(self):
        for _, c in getattr(self, "_lane_list", [])[1:]:
            c.close()
        self.ctx.close()

    # ---------------------------------------------------------------- per-clip image stages
    def fragments(self, frames, nexts, keep_intermediates=False, ctx=None):
        """A1-A8 for B pairs of one resolution: -> (ori_frag, merged_frag) [B,224,224,3] u8 BGR."""
        ctx = ctx or self.ctx
        r = ops.absdiff_patchsum(frames, nexts, want_residual=False, want_gray=True)
        pos, cnt = ops.topk_patches(r["sums"])
        ori, diff = ops.gather_fragments(frames, nexts, pos, cnt)
        flow, fsums, minmax = ops.farneback_flow_sums(ctx, r["gray0"], r["gray1"])
        fpos, fcnt = ops.topk_patches(fsums)
        flow_frag, merged = ops.flow_fragment_merge(flow, minmax, fpos, fcnt, diff, want_flow_frag=keep_intermediates)
        if keep_intermediates:
            return dict(ori_frag=ori, diff_frag=diff, merged_frag=merged, flow_frag=flow_frag, flow=flow, positions=pos,
                        count=cnt, flow_positions=fpos, sums=r["sums"], flow_sums=fsums)
        return ori, merged

    # --------------------------------------------------------------------------- full path
    #: clips whose image stages run side by side (each lane = one CUDA stream + one library context for its workspaces);
    #: the coarse pyramid levels launch grids far smaller than the GPU, which neighbouring clips fill
    LANES = 4

    #: False = everything on the current stream, one clip after the other (clean per-kernel timings for profiling)
    concurrent = True

    #: True = software pipeline across calls: the image stages (HBM-bound) of a call depend only on their clips being
    #: ready, not on the previous call's backbones (tensor-bound), so the lanes run one call ahead of the backbone streams
    #: and the two phases share the GPU instead of alternating.  False = every call starts after the previous one ended.
    #: Measured on the 1080p workload (profiles/r2_overlap_sweep.log): 93.0 videos/s piped vs 97.6 un-piped, for every
    #: gemm_sms setting - both phases are bound by per-SM resources (issue slots / shared memory / tensor pipe), so sharing
    #: the SMs adds no throughput and the interleaving costs the persistent GEMM grids their wave structure.  Off by default.
    pipeline = False

    #: SMs the persistent tcgen05 grids may occupy (0 = all 148); the rest stays free for the lanes' kernels
    gemm_sms = 0

    def set_gemm_sms(self, sms):
        self.gemm_sms = int(sms)
        self.ctx.set_gemm_sms(self.gemm_sms)

    def flow_kernel_name(self):
        return "k4_flow_iter_march"

    @property
    def launches(self):
        """Kernels launched by this engine (all lanes)."""
        return sum(c.launches for c in self._contexts())

    def _contexts(self):
        return [self.ctx] + [c for _, c in getattr(self, "_lane_list", [])[1:]]

    def set_profiling(self, on):
        for c in self._contexts():
            c.set_profiling(on)

    def profile_read_flow(self):
        """-> (ms, launches, algorithmic bytes) of the Farneback iteration launches of all lanes since the last read."""
        parts = [c.profile_read_flow() for c in self._contexts()]
        return tuple(sum(p[i] for p in parts) for i in range(3))

    def _lanes(self):
        if not hasattr(self, "_lane_list"):
            n = max(1, int(self.LANES))
            self._lane_list = [(torch.cuda.Stream(self.device), self.ctx if i == 0 else ops.Context(self.device.index))
                               for i in range(n)]
            self._side = torch.cuda.Stream(self.device)
        return self._lane_list

    #: clips of the same resolution are fused into one image-stage call of up to this many pair-pixels (0 = one call per clip)
    GROUP_PIXELS = 0

    def _units(self, clips):
        """-> lists of clip indices; every list is one image-stage call on one lane."""
        if not self.GROUP_PIXELS or not self.concurrent:
            return [[i] for i in range(len(clips))]
        units, open_units = [], {}
        for i, c in enumerate(clips):
            key = tuple(c.frames.shape[1:3])
            px = max(1, c.nexts.shape[0]) * key[0] * key[1]
            u = open_units.get(key)
            if u is not None and u[1] + px <= self.GROUP_PIXELS:
                u[0].append(i); u[1] += px
            else:
                u = [[i], px]
                open_units[key] = u
                units.append(u)
        return [u[0] for u in units]

    def extract_blocks(self, clips: Sequence[Clip]):
        """-> dict of per-frame feature matrices stacked over clips + row offsets per clip."""
        ctx = self.ctx
        main = torch.cuda.current_stream(self.device)
        lanes = self._lanes() if self.concurrent else [(main, ctx)]
        full_rn, full_vt, oris, mers = [], [], [], []
        full_off, pair_off = [0], [0]
        piped = self.concurrent and self.pipeline
        if not piped:
            for s, _ in lanes[:len(clips)]:
                s.wait_stream(main)
        for c in clips:
            if piped and c.ready is None:
                # first use of this clip: everything queued on the current stream so far (its producer) is the dependency;
                # later calls with the same (immutable) clip find the event complete and do not wait for earlier backbones
                c.ready = torch.cuda.Event()
                c.ready.record(main)
        n = len(clips)
        oris, mers, full_rn, full_vt = [None] * n, [None] * n, [None] * n, [None] * n
        for u, unit in enumerate(self._units(clips)):
            s, lctx = lanes[u % len(lanes)]
            with torch.cuda.stream(s):
                for i in unit:
                    if clips[i].ready is not None:
                        s.wait_event(clips[i].ready)
                tps = [clips[i].nexts.shape[0] for i in unit]
                tfs = [clips[i].frames.shape[0] for i in unit]
                if len(unit) == 1:
                    c = clips[unit[0]]
                    fr_pairs, nx_pairs, fr_all = c.frames[:tps[0]], c.nexts, c.frames
                else:       # clips of one resolution fused into one call: larger grids, fewer launches (results are batch-invariant)
                    fr_pairs = torch.cat([clips[i].frames[:tp] for i, tp in zip(unit, tps)])
                    nx_pairs = torch.cat([clips[i].nexts for i in unit])
                    fr_all = torch.cat([clips[i].frames for i in unit])
                ori, mer = self.fragments(fr_pairs, nx_pairs, ctx=lctx)
                rn, vt = ops.resize_pil_pair(lctx, fr_all)       # BILINEAR for the ResNet, LANCZOS for the ViT
            for t in [ori, mer, rn, vt] + [clips[i].frames for i in unit] + [clips[i].nexts for i in unit]:
                t.record_stream(s)
                t.record_stream(main)
            po = fo = 0
            for i, tp, tf in zip(unit, tps, tfs):
                oris[i], mers[i] = ori[po:po + tp], mer[po:po + tp]
                full_rn[i], full_vt[i] = rn[fo:fo + tf], vt[fo:fo + tf]
                po += tp; fo += tf
        for c in clips:
            full_off.append(full_off[-1] + c.frames.shape[0])
            pair_off.append(pair_off[-1] + c.nexts.shape[0])
        for s, _ in lanes[:len(clips)]:
            main.wait_stream(s)
        nf, npair = full_off[-1], pair_off[-1]
        rn_in = torch.cat(full_rn + oris + mers)
        vt_in = torch.cat(full_vt + oris + mers)
        # the two backbones on two streams: each launch is a persistent grid over all SMs, so the blocks of one network's
        # next kernel start on the SMs the other network's last wave has already left
        side = self._side if self.concurrent else main
        side.wait_stream(main)
        with torch.cuda.stream(side):
            vit = ops.vitb16_features(ctx, vt_in, is_bgr=True)
        vt_in.record_stream(side)
        vit.record_stream(main)
        stack, pool = ops.resnet50_features(ctx, rn_in, is_bgr=True, want_stack=True, want_pool=True)
        main.wait_stream(side)
        dev = self.device
        return dict(full_resnet=stack[:nf], full_vit=vit[:nf], frag_stack=stack[nf:nf + npair],
                    frag_pool=pool[nf + npair:], frag_vit_ori=vit[nf:nf + npair], frag_vit_mer=vit[nf + npair:],
                    full_off=self._offsets(tuple(full_off)), pair_off=self._offsets(tuple(pair_off)))

    def _offsets(self, key):
        """Row offsets per clip as a device int32 tensor.  Cached per distinct tuple: a pageable host-to-device copy
        synchronises the host with the stream, which would stop it from queueing the next batch behind this one."""
        cache = self.__dict__.setdefault("_offset_cache", {})
        t = cache.get(key)
        if t is None:
            if len(cache) > 4096:
                cache.clear()
            t = cache[key] = torch.tensor(key, dtype=torch.int32, device=self.device)
        return t

    def extract(self, clips: Sequence[Clip]):
        """-> features [V, 35203] fp32 on the device (layout of src/demo_test.py:171-175)."""
        b = self.extract_blocks(clips)
        return ops.temporal_mean_concat(b["full_resnet"].contiguous(), b["full_vit"].contiguous(), b["frag_stack"].contiguous(),
                                        b["frag_pool"].contiguous(), b["frag_vit_ori"].contiguous(), b["frag_vit_mer"].contiguous(),
                                        b["full_off"], b["pair_off"])

    def predict(self, clips: Sequence[Clip], video_type: Optional[str] = None, is_finetune=False):
        """-> (features [V,35203], scores [V]); scores rescaled like src/demo_test.py:206-219."""
        if not self.has_head:
            raise ops._lib.B200VQAError("no regression head loaded")
        feats = self.extract(clips)
        score = ops.head_forward(self.ctx, feats)
        if not is_finetune and video_type in ("youtube_ugc", "konvid_1k"):
            score = (score / 100.0) * 4.0 + 1.0
        return feats, score

    def submit_host(self, host_clips: Sequence[Sequence[torch.Tensor]], video_type=None):
        """Asynchronous end-to-end entry with HOST (pinned) uint8 buffers: queues the H2D copies on a side stream (one
        event per clip, so the copy of clip i+1 overlaps the fragment stages of clip i), the full path on the current
        stream and the D2H copy of the scores into a pinned buffer.  Returns a ticket for `result()` (collect it before
        eight further submits).  Submitting
        batch k+1 before collecting batch k lets its copies run under batch k's kernels."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)
        clips = []
        for f, n in host_clips:
            with torch.cuda.stream(self._copy_stream):
                # blocks come from the copy stream's pool; record_stream defers their reuse until the kernels are done
                df = f.to(self.device, non_blocking=True)
                dn = n.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            df.record_stream(main)
            dn.record_stream(main)
            clips.append(Clip(df, dn, ev))
        feats, score = self.predict(clips, video_type)
        # pinned score buffers are taken from a small ring owned by the engine: no pinned-memory allocator call (which can
        # synchronise with the device) on the per-batch path; result() copies the scores out before the slot comes round again
        ring = self.__dict__.setdefault("_score_ring", [])
        n = score.numel()
        if not ring or ring[0].numel() < n:
            ring[:] = [torch.empty(max(n, 64), dtype=score.dtype, pin_memory=True) for _ in range(8)]
            self._score_slot = 0
        self._score_slot = (self._score_slot + 1) % len(ring)
        host_score = ring[self._score_slot][:n]
        host_score.copy_(score, non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        return feats, host_score, done

    @staticmethod
    def result(ticket):
        """Waits for a `submit_host` ticket -> (features on the device, scores on the host)."""
        feats, host_score, done = ticket
        done.synchronize()
        return feats, host_score.clone()

    def predict_host(self, host_clips: Sequence[Sequence[torch.Tensor]], video_type=None):
        """End-to-end entry with HOST (pinned) uint8 buffers: H2D copies, full path, D2H of the scores."""
        return self.result(self.submit_host(host_clips, video_type))


def synthetic_clips_on_device(n_clips, H, W, pairs, device, seed=0) -> List[Clip]:
    """Device-side generator for benchmark inputs (bulk data; parity tests use synth.make_clip on the
    host instead).  Blurred-noise texture, global shift, one moved block, +-3 noise."""
    g = torch.Generator(device=device).manual_seed(seed)
    k = torch.tensor([1, 4, 7, 10, 7, 4, 1], dtype=torch.float32, device=device)
    k = (k / k.sum())
    clips = []
    for _ in range(n_clips):
        noise = torch.rand((pairs, 3, H + 32, W + 32), generator=g, device=device) * 255.0
        base = noise
        for _ in range(3):   # repeated separable binomial blur ~ sigma 3.3
            base = torch.nn.functional.conv2d(base.reshape(-1, 1, H + 32, W + 32), k.view(1, 1, 1, 7), padding=(0, 3))
            base = torch.nn.functional.conv2d(base, k.view(1, 1, 7, 1), padding=(3, 0)).reshape(pairs, 3, H + 32, W + 32)
        base = (base - base.mean()) * 6.0 + 128.0
        frames = base[:, :, 16:16 + H, 16:16 + W]
        dx, dy = (int(v) for v in torch.randint(-3, 4, (2,), generator=g, device=device).tolist())
        moved = torch.roll(base, shifts=(dy, dx), dims=(2, 3)) * 0.5 + torch.roll(base, shifts=(dy, dx + 1), dims=(2, 3)) * 0.5
        bh, bw = min(130, H // 3), min(160, W // 3)
        by, bx = 16 + H // 4, 16 + W // 4
        moved[:, :, by + 5:by + 5 + bh, bx + 10:bx + 10 + bw] = base[:, :, by:by + bh, bx:bx + bw]
        moved = moved + (torch.rand(moved.shape, generator=g, device=device) * 6.0 - 3.0)
        nexts = moved[:, :, 16:16 + H, 16:16 + W]
        to_u8 = lambda t: t.clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
        clips.append(Clip(to_u8(frames), to_u8(nexts)))
    return clips
