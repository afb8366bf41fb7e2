"""Dataset-scale extraction job: the ``__main__`` loops of the reference's
``main_layer_stack.py`` / ``main_fragment_layerstack.py`` / ``main_fragment_pool.py`` (ref
src/main_fragment_layerstack.py:262-364) as one batched, restartable, video-sharded job over pre-sampled frames.

Input : a metadata CSV with the reference's columns (vid, width, height, framerate, nb_frames, ...;
        metadata/*.csv) and, per video, the PNGs the reference's sampler writes
        (``{vid}_{i}.png`` / ``{vid}_{i}_next.png`` in ``<frames_root>/video_{i+1}/``).
Output: per video four ``.npy`` files with the reference's names and shapes:
        resnet50/layer_stack (T,13120) | vit/pool (T,2304) | fragment resnet50/layer_stack (T,15171) |
        fragment vit/pool (T,4608).

* resume: a video whose four files all exist AND load is skipped; files are written to a temporary name and
  renamed, so a killed job never leaves a truncated file that later passes for done;
* greyscale filter: rows listed in the reference's greyscale report (first column = row index,
  src/data_processing/split_train_test.py:113-117) are not extracted (status "greyscale");
* sharding: with world > 1 every rank computes the same longest-processing-time plan from the metadata
  (pairs x resolution cost model, ``sharding.shard_videos``) and extracts its own videos - the outputs are per-video
  files, so ranks exchange nothing but the status list at the end (``all_gather_object`` when a process group exists);
* loading: PNG decoding of the next batch runs on a host thread while the GPU works on the current one;
* a video that cannot be read gets a status entry instead of aborting the batch."""
import math
import os
import threading
import queue

import numpy as np
import torch

from . import sharding
from .data_processing import extract_npy2mat as fmt
from .demo_test import load_clip_host
from .engine import Clip, Engine


def frame_interval(framerate):
    """src/main_fragment_layerstack.py:274-277."""
    return math.ceil(framerate / 2) if framerate < 2 else int(framerate / 2)


def sampled_pairs(nb_frames, framerate):
    k = max(1, frame_interval(float(framerate)))
    n_full = -(-int(nb_frames) // k)
    n_next = -(-(int(nb_frames) - 1) // k)
    return max(0, min(n_full, n_next))


def output_paths(out_root, data_name, index, resolution=None):
    if data_name == "youtube_ugc" and not resolution:
        raise ValueError("youtube_ugc needs `resolution` (e.g. '360P'): the reference keeps one feature folder and one "
                         "file suffix per resolution (src/data_processing/extract_npy2mat.py:55-60)")
    d = lambda base, net, layer: fmt.features_dir(os.path.join(out_root, base), net, layer, data_name, resolution=resolution)
    n = lambda net: fmt.npy_name(index, net, resolution=resolution if data_name == "youtube_ugc" else None)
    return {
        "full_resnet": os.path.join(d("features", "resnet50", "layer_stack"), n("resnet50")),
        "full_vit": os.path.join(d("features", "vit", "pool"), n("vit")),
        "frag_resnet": os.path.join(d("features_merged_frag", "resnet50", "layer_stack"), n("resnet50")),
        "frag_vit": os.path.join(d("features_merged_frag", "vit", "pool"), n("vit")),
    }


def atomic_save_npy(path, array):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        np.save(f, array)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)


def _is_complete(path):
    """The file exists and its header + payload size agree (a truncated file from a killed writer is not "done")."""
    try:
        a = np.load(path, mmap_mode="r")
        return a.ndim == 2 and a.shape[0] >= 0
    except Exception:
        return False


def greyscale_indices(greyscale_csv):
    """Row indices to drop: first column of the reference's *_greyscale_metadata.csv (split_train_test.py:113-117)."""
    import pandas as pd
    return set(int(i) for i in pd.read_csv(greyscale_csv).iloc[:, 0].tolist())


def run(metadata_csv, frames_root, out_root, data_name, engine: Engine = None, batch_videos=4, limit=None, resolution=None,
        greyscale_csv=None, rank=None, world=None, prefetch=2):
    """-> list of (vid, status), status in {"done", "skipped", "greyscale", "error: ..."}; with world > 1 and an initialised
    process group the list covers all ranks (in metadata order), otherwise this rank's videos."""
    import pandas as pd
    import torch.distributed as dist
    meta = pd.read_csv(metadata_csv)
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if engine is None:
        from . import runtime
        engine = runtime.engine()
    eng = engine
    rows = list(meta.itertuples(index=True))[:limit]
    grey = greyscale_indices(greyscale_csv) if greyscale_csv else set()
    has_geo = all(hasattr(rows[0], c) for c in ("width", "height", "framerate", "nb_frames")) if rows else False
    costs = [sharding.video_cost(sampled_pairs(r.nb_frames, r.framerate), int(r.height), int(r.width)) if has_geo else 1.0 for r in rows]
    mine = set(sharding.shard_videos(costs, world)[rank])
    status = {}

    todo = []
    for pos, r in enumerate(rows):
        if pos not in mine:
            continue
        idx, vid = r.Index, str(r.vid)
        if idx in grey:
            status[pos] = (vid, "greyscale")
            continue
        paths = output_paths(out_root, data_name, idx, resolution)
        if all(_is_complete(p) for p in paths.values()):
            status[pos] = (vid, "skipped")
            continue
        todo.append((pos, idx, vid))

    # ---- host thread: decode the PNGs of upcoming videos into pinned memory while the GPU works
    q = queue.Queue(maxsize=max(1, prefetch) * batch_videos)

    def loader():
        for pos, idx, vid in todo:
            folder = os.path.join(frames_root, f"video_{idx + 1}")
            try:
                frames, nexts = load_clip_host(folder, folder, vid, pin=True)
                if nexts.shape[0] == 0:
                    raise ValueError("no sampled pairs")
                q.put((pos, idx, vid, frames, nexts, None))
            except Exception as e:          # keep going: the reference would abort the whole run here
                q.put((pos, idx, vid, None, None, f"error: {e}"))
        q.put(None)

    t = threading.Thread(target=loader, daemon=True)
    t.start()
    pending = []

    def flush():
        if not pending:
            return
        clips = [Clip(f.to(eng.device, non_blocking=True), n.to(eng.device, non_blocking=True)) for _, _, _, f, n in pending]
        blocks = eng.extract_blocks(clips)
        fo, po = blocks["full_off"].tolist(), blocks["pair_off"].tolist()
        for k, (pos, idx, vid, _f, _n) in enumerate(pending):
            mats = {
                "full_resnet": blocks["full_resnet"][fo[k]:fo[k + 1]],
                "full_vit": blocks["full_vit"][fo[k]:fo[k + 1]],
                "frag_resnet": torch.cat([blocks["frag_stack"][po[k]:po[k + 1]], blocks["frag_pool"][po[k]:po[k + 1]]], dim=1),
                "frag_vit": torch.cat([blocks["frag_vit_ori"][po[k]:po[k + 1]], blocks["frag_vit_mer"][po[k]:po[k + 1]]], dim=1),
            }
            for key, path in output_paths(out_root, data_name, idx, resolution).items():
                atomic_save_npy(path, mats[key].cpu().numpy())
            status[pos] = (vid, "done")
        pending.clear()

    while True:
        item = q.get()
        if item is None:
            break
        pos, idx, vid, frames, nexts, err = item
        if err:
            status[pos] = (vid, err)
            continue
        # clips of one batch may differ in resolution: the fragment stages run per clip, the backbones on the union
        pending.append((pos, idx, vid, frames, nexts))
        if len(pending) >= batch_videos:
            flush()
    flush()
    t.join()
    if world > 1 and dist.is_initialized():
        parts = [None] * world
        dist.all_gather_object(parts, status)
        status = {k: v for p in parts for k, v in p.items()}
    return [status[k] for k in sorted(status)]
