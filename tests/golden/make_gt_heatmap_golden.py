"""Writes tests/golden/ref_gt_heatmaps.npz from the UNMODIFIED reference loader (dataloader/data_loader.py
process_frame_data + the wrapper's channel concatenation), run in the build container where /root/reference exists.
scikit-image is absent there: skimage.draw.line_aa is served by the restated algorithm in oracle/gt_heatmap_oracle.py
(see that file's header) -- everything else is the reference's own code.

    python tests/golden/make_gt_heatmap_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import gt_heatmap_oracle as gto  # noqa: E402
import test_gt_heatmap_oracle as T  # noqa: E402


def main():
    T._reference_modules()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for preset in ("UnrealEgo", "EgoCap"):
            pts2d, pts3d = gto.synthetic_keypoints(preset, 1, seed=11)
            p3 = np.stack([pts3d[0, 0], pts3d[0, 0] + (pts3d[0, 1][0] - pts3d[0, 0][0])[None]])[None]
            out[preset + "_pts2d"] = pts2d
            out[preset + "_pts3d"] = p3.astype(np.float32)
            out[preset + "_input"] = T._reference_lifting_input(preset, pts2d[0], p3[0], tmp)[None].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_gt_heatmaps.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
