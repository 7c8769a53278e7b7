"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the ground-truth heatmap synthesis that feeds the lifting network when the
reference runs with ``--use_gt_heatmap``: from 2-D / 3-D keypoints to the
(6J, 64, 64) stereo joint + limb heatmap stack ``[joint L | joint R | cos L | sin L | cos R | sin R]``.

Restates, in numpy:
  * joint heatmaps  -- reference utils/projection.py:263-279 (coord2d_to_heatmap): a unit impulse at the truncated pixel
    position on a (res + 8)^2 canvas, scipy gaussian_filter (sigma 1, truncate 4 -> 9 taps), cropped, divided by
    1 / (2 pi) = 0.15915589174187972
  * limb heatmaps   -- reference utils/data.py:175-185, 197-252 (get_line_limb_heatmap, get_limb_data): endpoints rounded to
    pixels (np.rint), anti-aliased line, gaussian_filter(sigma 1, mode='constant'), * sigma
  * sin/cos modulation and stacking -- reference dataloader/data_loader.py:127-132, 193-199 (x2, cos / sin of the limb's
    elevation angle, taken from the LEFT view's 3-D points for both views, utils/data.py:254-262) and
    model/egotap_autoencoder_model.py:176-181, 199-213 (channel order of pred_heatmap_cat)

Third-party arithmetic that is NOT in the reference tree:
  * scipy.ndimage.gaussian_filter (reference pins scipy==1.7.3, requirements.txt:70): separable correlation with the
    normalised weights exp(-k^2 / 2 sigma^2), radius int(4 sigma + 0.5), boundary 'reflect' (default) or 'constant'.
    scipy IS installed here (newer version): ``gaussian_blur`` below is checked against it live in the tests.
  * skimage.draw.line_aa (reference pins scikit-image==0.19.3, requirements.txt:68) is NOT installed and cannot be
    (no network).  ``line_aa`` below restates its published algorithm (skimage/draw/_draw.pyx ``_line_aa``: Zingl's
    anti-aliased Bresenham line) and is pinned only to the known-answer example in skimage's own docstring
    (line_aa(1, 1, 8, 8): 255 on the diagonal, 74 beside it) -- "parity unpinned" beyond that; the test that compares
    with the reference's get_limb_data injects this same restatement for the missing import.
"""
import math

import numpy as np

RES = 64
INV_2PI = 0.15915589174187972
KINEMATIC_PARENTS = {   # reference utils/util.py:51-52
    "UnrealEgo": [0, 0, 1, 1, 2, 3, 4, 5, 2, 3, 8, 9, 10, 11, 12, 13],
    "EgoCap": [0, 0, 1, 2, 3, 4, 1, 6, 7, 8, 2, 10, 11, 12, 6, 14, 15, 16],
}


def gaussian_weights(sigma=1.0, truncate=4.0):
    r = int(truncate * sigma + 0.5)
    k = np.arange(-r, r + 1, dtype=np.float64)
    w = np.exp(-0.5 * k * k / (sigma * sigma))
    return w / w.sum()


def gaussian_blur(img, sigma=1.0, mode="reflect"):
    """scipy.ndimage.gaussian_filter for a 2-D float32 image: axis 0 then axis 1, float32 between the passes"""
    w = gaussian_weights(sigma)
    r = len(w) // 2
    out = img.astype(np.float32)
    for axis in (0, 1):
        a = np.moveaxis(out, axis, 0).astype(np.float64)
        pad = np.pad(a, ((r, r), (0, 0)), mode="symmetric" if mode == "reflect" else "constant")
        acc = np.zeros_like(a)
        for k in range(2 * r + 1):
            acc += w[k] * pad[k:k + a.shape[0]]
        out = np.moveaxis(acc, 0, axis).astype(np.float32)
    return out


def coord2d_to_heatmap(coord2d, res=RES, sigma=1.0):
    """reference utils/projection.py:263-279.  coord2d: (n, 2) in 1024-pixel image coordinates (x, y)."""
    hm = np.zeros((coord2d.shape[0], res, res), dtype=np.float32)
    m = int(4 * sigma)
    for i in range(coord2d.shape[0]):
        x, y = coord2d[i] / 1024.0 * res
        canvas = np.zeros((res + 2 * m, res + 2 * m), dtype=np.float32)
        if -4 <= y < res + 4 and -4 <= x < res:          # (sic) the x test has no +4: specification
            canvas[int(y) + m, int(x) + m] = 1.0
        hm[i] = gaussian_blur(canvas, sigma)[m:-m, m:-m]
    return hm / np.float32(INV_2PI)


def line_aa(r0, c0, r1, c1):
    """skimage.draw.line_aa (scikit-image 0.19.3, skimage/draw/_draw.pyx:_line_aa) restated: returns (rr, cc, val)"""
    rr, cc, val = [], [], []
    dc, dr = abs(c0 - c1), abs(r0 - r1)
    err = np.float32(dc - dr)                            # `cdef float err`: single precision in the original
    sign_c = 1 if c0 < c1 else -1
    sign_r = 1 if r0 < r1 else -1
    ed = np.float32(1.0) if dc + dr == 0 else np.float32(math.sqrt(dc * dc + dr * dr))
    c, r = c0, r0
    while True:
        cc.append(c); rr.append(r)
        val.append(abs(float(err) - dc + dr) / float(ed))
        err_prime, c_prime = err, c
        if 2 * float(err_prime) >= -dc:
            if c == c1:
                break
            if float(err_prime) + dr < float(ed):
                cc.append(c); rr.append(r + sign_r)
                val.append(abs(float(err_prime) + dr) / float(ed))
            err = np.float32(float(err) - dr)
            c += sign_c
        if 2 * float(err_prime) <= dr:
            if r == r1:
                break
            if dc - float(err_prime) < float(ed):
                cc.append(c_prime + sign_c); rr.append(r)
                val.append(abs(dc - float(err_prime)) / float(ed))
            err = np.float32(float(err) + dc)
            r += sign_r
    return np.array(rr, dtype=np.intp), np.array(cc, dtype=np.intp), 1.0 - np.array(val, dtype=np.float64)


def line_limb_heatmap(p_coord, coord, res=RES):
    """reference utils/data.py:175-185: pixels are ASSIGNED in generation order (a later duplicate overwrites)"""
    hm = np.zeros((res, res), dtype=np.float32)
    p = np.rint(p_coord).astype(int)
    c = np.rint(coord).astype(int)
    rr, cc, val = line_aa(p[0], p[1], c[0], c[1])
    for x, y, v in zip(rr, cc, val):
        if 0 <= x <= res - 1 and 0 <= y <= res - 1:
            hm[y, x] = v
    return hm


def limb_data(pts2d, pts3d, preset, res=RES, sigma=1.0):
    """reference utils/data.py:197-252 (htype 'line', area == res): (J, res, res) raw limb heatmaps and theta (J,)"""
    parents = KINEMATIC_PARENTS[preset]
    n = len(parents)
    hms = np.zeros((n - 1, res, res), dtype=np.float32)
    theta = np.zeros(n - 1, dtype=np.float32)
    for j in range(1, n):
        q = parents[j]
        l3 = pts3d[q] - pts3d[j]
        with np.errstate(divide="ignore", invalid="ignore"):
            theta[j - 1] = np.arctan(l3[2] / np.linalg.norm(l3[:2]))
        hm = line_limb_heatmap(pts2d[q] / (1024.0 / res), pts2d[j] / (1024.0 / res), res)
        hms[j - 1] = gaussian_blur(hm, sigma, mode="constant") * sigma
    return hms, theta


def lifting_input(pts2d_left, pts2d_right, pts3d_left, pts3d_right, preset):
    """one frame's (6J, 64, 64) input of the lifting network from keypoints, as the reference's loader + wrapper build it
    with --use_gt_heatmap (dataloader/data_loader.py:76-208, model/egotap_autoencoder_model.py:176-213)"""
    joint_l = coord2d_to_heatmap(pts2d_left[1:])
    joint_r = coord2d_to_heatmap(pts2d_right[1:])
    limb_l, theta = limb_data(pts2d_left, pts3d_left, preset)
    limb_r, _ = limb_data(pts2d_right, pts3d_right, preset)          # theta of the LEFT view is used for both
    cos, sin = np.cos(theta.astype(np.float32))[:, None, None], np.sin(theta.astype(np.float32))[:, None, None]
    limb_l, limb_r = limb_l * 2, limb_r * 2
    return np.concatenate([joint_l, joint_r, limb_l * cos, limb_l * sin, limb_r * cos, limb_r * sin], 0).astype(np.float32)


def synthetic_keypoints(preset, batch, seed=0):
    """plausible keypoints for tests / benchmarks: 2-D points over (and slightly outside) the 1024-pixel fisheye image,
    3-D points in centimetres.  Returns pts2d (B, 2, n, 2) float32 and pts3d (B, 2, n, 3) float32 (view-specific: local
    pose + that view's pelvis offset, as dataloader/data_loader.py:107-115)."""
    rng = np.random.default_rng(seed)
    n = len(KINEMATIC_PARENTS[preset])
    pts2d = rng.uniform(-80, 1100, size=(batch, 2, n, 2)).astype(np.float32)
    local = rng.normal(0, 30, size=(batch, 1, n, 3)).astype(np.float32)
    pelvis = rng.normal(0, 10, size=(batch, 2, 1, 3)).astype(np.float32)
    return pts2d, (local + pelvis).astype(np.float32)
