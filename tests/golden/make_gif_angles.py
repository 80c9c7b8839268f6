"""Joint angles of the 2-link arm read back from the animations the reference ships — outputs of iLQR.jl itself
(test/2_link_example/animate_2_link.jl:27-41 draws x̄ᶠ[t, 1:2] of `iLQR.fit` every 10th knot point, 91 frames for
num_steps = 900, into test/2_link_example/figures/*.gif).  They are the only numbers of the reference's own runs that
exist anywhere, so they pin the CPU oracle against the real thing — to pixel accuracy (≈ 0.01 rad).

    python tests/golden/make_gif_angles.py  →  tests/golden/reference_gif_angles.json        (needs /root/reference)

(docs/extras/iLQR_2_link.gif is not used: it starts from θ = (π/4, −π/2) and ends in the mirrored elbow configuration of
the target (0.6, 0), which the joint-space cost of today's 2_link_helper_functions.jl cannot produce — it predates it.)

Method: axes from the five grid / frame lines of the first frame (x, y = −2 … 2 ⇒ origin and pixels per unit), the arm from its colour
(the only saturated pixels of a frame); per link a total-least-squares line through the pixels within 3.5 px of the link,
started from the marker centroids (darker where line and markers overlap), iterated twice."""
import json
import os
import sys

import numpy as np
from PIL import Image

FIG = "/root/reference/test/2_link_example/figures"
L1 = L2 = np.sqrt(2.0) / 2.0      # 2_link_helper_functions.jl:5


def axes_of(a):
    """pixel of the origin and pixels per unit from the five grid / frame lines x, y = −2 … 2 (xlims!, ylims!: ±2)"""
    s = a.sum(axis=2)
    gray = ((a.max(axis=2) - a.min(axis=2)) < 12) & (s < 760)        # neither white nor coloured

    def lines(counts):
        idx = [i for i, v in enumerate(counts) if v > 150]
        groups = [[idx[0]]]
        for i in idx[1:]:
            if i - groups[-1][-1] <= 2:
                groups[-1].append(i)
            else:
                groups.append([i])
        c = [float(np.mean(g)) for g in groups]
        assert len(c) == 5, c
        return c

    cx, cy = lines(gray.sum(axis=0)), lines(gray.sum(axis=1))
    return cx[2], cy[2], (cx[4] - cx[0]) / 4.0, (cy[4] - cy[0]) / 4.0


def fit_link(px, py, p0, p1):
    """direction (unit, from p0 towards p1) of the TLS line through the mask pixels near segment p0-p1"""
    d = p1 - p0
    L = np.hypot(*d); d = d / L
    rel = np.stack([px - p0[0], py - p0[1]], axis=1)
    t = rel @ d
    perp = np.abs(rel @ np.array([-d[1], d[0]]))
    sel = (t > 0.1 * L) & (t < 0.9 * L) & (perp < 3.5)
    pts = np.stack([px[sel], py[sel]], axis=1)
    c = pts.mean(axis=0)
    _, _, vt = np.linalg.svd(pts - c)
    v = vt[0]
    if v @ d < 0:
        v = -v
    return v, c


def angles_of(path):
    im = Image.open(path)
    out = []
    ax = None
    for f in range(im.n_frames):
        im.seek(f)
        a = np.array(im.convert("RGB")).astype(int)
        if ax is None:
            ax = axes_of(a)
        cx, cy, sx, sy = ax
        sat = a.max(axis=2) - a.min(axis=2)
        ys, xs = np.nonzero(sat > 60)
        px, py = xs.astype(float), ys.astype(float)
        o = np.array([cx, cy])
        mask = sat > 60
        Hh, Ww = mask.shape
        asp = np.array([1.0, sy / sx])
        l1px, l2px = L1 * sx, L2 * sx

        def sweep(p0, lpx, exclude=None):
            """coarse direction of the link that starts at p0: the angle whose ray lies on arm pixels the longest"""
            best, best_sc = None, -1.0
            for th in np.linspace(-np.pi, np.pi, 1440, endpoint=False):
                if exclude is not None and abs((th - exclude + np.pi) % (2 * np.pi) - np.pi) < 0.35:
                    continue
                d = np.array([np.cos(th), -np.sin(th)]) * asp
                t = np.linspace(0.2, 0.95, 40)[:, None] * lpx
                q = np.rint(p0 + t * d).astype(int)
                ok = (q[:, 0] >= 0) & (q[:, 0] < Ww) & (q[:, 1] >= 0) & (q[:, 1] < Hh)
                sc = mask[q[ok, 1], q[ok, 0]].mean() if ok.any() else 0.0
                if sc > best_sc:
                    best, best_sc = th, sc
            return best

        th1c = sweep(o, l1px)
        e = o + np.array([np.cos(th1c), -np.sin(th1c)]) * asp * l1px
        th2c = sweep(e, l2px, exclude=(th1c + np.pi))
        tip = e + np.array([np.cos(th2c), -np.sin(th2c)]) * asp * l2px
        for _ in range(3):
            v1, _ = fit_link(px, py, o, e)
            e = o + v1 / np.hypot(v1[0], v1[1] * sx / sy) * l1px   # elbow from the fitted direction and the known link length
            v2, _ = fit_link(px, py, e, tip)
            tip = e + v2 / np.hypot(v2[0], v2[1] * sx / sy) * l2px
        th1 = np.arctan2(-v1[1] / sy, v1[0] / sx)          # pixel rows grow downwards
        th12 = np.arctan2(-v2[1] / sy, v2[0] / sx)
        out.append([float(th1), float(th12)])
    return out, ax


if __name__ == "__main__":
    res = {}
    for name in sorted(os.listdir(FIG)):
        if not name.startswith("iLQR_2_link"):
            continue
        ang, ax = angles_of(os.path.join(FIG, name))
        res[name] = {"frames": len(ang), "knot_stride": 10, "axes_px": ax, "theta1_theta12": ang}
        print(name, len(ang), "first", np.round(ang[0], 3), "last", np.round(ang[-1], 3))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_gif_angles.json")
    json.dump({"source": "/root/reference/test/2_link_example/figures/*.gif (animate_2_link.jl:27-41)",
               "columns": "[theta1, theta1 + theta2] per frame, radians, read from pixels (≈ ±0.01 rad)", "gifs": res}, open(dst, "w"))
    print("wrote", dst)
