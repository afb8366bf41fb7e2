"""Dataset-scale extraction job: the ``__main__`` loops of the reference's
``main_layer_stack.py`` / ``main_fragment_layerstack.py`` / ``main_fragment_pool.py`` as one batched,
restartable job over pre-sampled frames.

Input : a metadata CSV with the reference's columns (vid, width, height, framerate, nb_frames, ...;
        metadata/*.csv) and, per video, the PNGs the reference's sampler writes
        (``{vid}_{i}.png`` / ``{vid}_{i}_next.png`` in ``<frames_root>/video_{i+1}/``).
Output: per video four ``.npy`` files with the reference's names and shapes:
        resnet50/layer_stack (T,13120) | vit/pool (T,2304) | fragment resnet50/layer_stack (T,15171) |
        fragment vit/pool (T,4608); a video whose files all exist is skipped (resume).  A video that cannot be
        read gets a status entry instead of aborting the batch."""
import math
import os

import numpy as np
import torch

from .data_processing import extract_npy2mat as fmt
from .demo_test import load_clip
from .engine import Engine


def frame_interval(framerate):
    """src/main_fragment_layerstack.py:274-277."""
    return math.ceil(framerate / 2) if framerate < 2 else int(framerate / 2)


def output_paths(out_root, data_name, index):
    d = lambda base, net, layer: fmt.features_dir(os.path.join(out_root, base), net, layer, data_name)
    return {
        "full_resnet": os.path.join(d("features", "resnet50", "layer_stack"), fmt.npy_name(index, "resnet50")),
        "full_vit": os.path.join(d("features", "vit", "pool"), fmt.npy_name(index, "vit")),
        "frag_resnet": os.path.join(d("features_merged_frag", "resnet50", "layer_stack"), fmt.npy_name(index, "resnet50")),
        "frag_vit": os.path.join(d("features_merged_frag", "vit", "pool"), fmt.npy_name(index, "vit")),
    }


def run(metadata_csv, frames_root, out_root, data_name, engine: Engine = None, batch_videos=4, limit=None):
    """-> list of (vid, status) with status in {"done", "skipped", "error: ..."}."""
    import pandas as pd
    meta = pd.read_csv(metadata_csv)
    eng = engine or Engine(0)
    rows = list(meta.itertuples(index=True))[:limit]
    status, pending = [], []

    def flush():
        if not pending:
            return
        blocks = eng.extract_blocks([c for _, _, c in pending])
        fo, po = blocks["full_off"].tolist(), blocks["pair_off"].tolist()
        for k, (idx, vid, _clip) in enumerate(pending):
            mats = {
                "full_resnet": blocks["full_resnet"][fo[k]:fo[k + 1]],
                "full_vit": blocks["full_vit"][fo[k]:fo[k + 1]],
                "frag_resnet": torch.cat([blocks["frag_stack"][po[k]:po[k + 1]], blocks["frag_pool"][po[k]:po[k + 1]]], dim=1),
                "frag_vit": torch.cat([blocks["frag_vit_ori"][po[k]:po[k + 1]], blocks["frag_vit_mer"][po[k]:po[k + 1]]], dim=1),
            }
            for key, path in output_paths(out_root, data_name, idx).items():
                os.makedirs(os.path.dirname(path), exist_ok=True)
                np.save(path, mats[key].cpu().numpy())
            status.append((vid, "done"))
        pending.clear()

    for r in rows:
        idx, vid = r.Index, str(r.vid)
        paths = output_paths(out_root, data_name, idx)
        if all(os.path.exists(p) for p in paths.values()):
            status.append((vid, "skipped"))
            continue
        folder = os.path.join(frames_root, f"video_{idx + 1}")
        try:
            clip = load_clip(folder, folder, vid, eng.device)
            if clip.nexts.shape[0] == 0:
                raise ValueError("no sampled pairs")
        except Exception as e:          # keep going: the reference would abort the whole run here
            status.append((vid, f"error: {e}"))
            continue
        # clips of one batch may differ in resolution: the fragment stages run per clip, the backbones on the union
        pending.append((idx, vid, clip))
        if len(pending) >= batch_videos:
            flush()
    flush()
    return status
