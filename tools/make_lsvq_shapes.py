"""Builds relax_vqa_b200/data/lsvq_train_shapes.csv: the (width, height, sampled pairs) histogram of the reference's
metadata/LSVQ_TRAIN_metadata.csv (28,013 rows).  Run in the build container only (needs
/root/reference); the table is what `bench.py --workload lsvq-mix` draws its mixed-resolution clips from."""
import math
import os
import sys

import pandas as pd

REF = os.environ.get("RELAXVQA_REF", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "relax_vqa_b200", "data", "lsvq_train_shapes.csv")


def pairs(nb, fps):
    k = max(1, math.ceil(fps / 2) if fps < 2 else int(fps / 2))          # src/main_fragment_layerstack.py:274-277
    return min(-(-int(nb) // k), -(-(int(nb) - 1) // k))                  # the two ffmpeg select filters, vf_extract.py:51-74


def main():
    df = pd.read_csv(os.path.join(REF, "metadata", "LSVQ_TRAIN_metadata.csv"))
    df["pairs"] = [pairs(a, b) for a, b in zip(df.nb_frames, df.framerate)]
    g = df.groupby(["width", "height", "pairs"]).size().reset_index(name="n").sort_values(["n", "width", "height", "pairs"],
                                                                                          ascending=[False, True, True, True])
    g.to_csv(OUT, index=False)
    print(OUT, len(g), "rows;", int(g.n.sum()), "videos;", int((g.width * g.height * g.pairs * g.n).sum()), "pair-pixels")


if __name__ == "__main__":
    sys.exit(main())
