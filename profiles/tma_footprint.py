#!/usr/bin/env python
"""Would a TMA box (cp.async.bulk.tensor, 4-D [x][y][z][record]) around a CTA's marching window be mostly useful bytes?

For every sampled CTA tile (16x8 pixels = 128 rays, the forward kernels' CTA) of the c3 workload (256^3 grid, 800x800,
256 jittered samples/ray) and K consecutive marching steps, this counts
    touched = distinct voxels referenced as trilinear corners by the tile's in-grid samples of those K steps
    box     = voxels of the axis-aligned bounding box of those corners (what one TMA box copy would have to move)
and prints the distribution of touched / box (the fraction of a box copy that any sample of the window reads) plus the
box size in bytes at 112-byte records.  Pure geometry (NumPy, CPU): positions follow sample.py:54-67 / voxels.py:214-223.

    python profiles/tma_footprint.py > profiles/r02_tma_footprint.json
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
from cases import HOTDOG_RADIUS, spherical_pose  # noqa: E402

G, SIDE, S, NEAR, FAR, WORLD = 256, 800, 256, 1.8, 6.6, 3.0
REC = 112


def tile_stats(rot, trans, focal, tiles, ks, rng, tile_w=16, tile_h=8):
    out = {k: {"ratio": [], "box_voxels": [], "touched": []} for k in ks}
    t = np.linspace(0.0, 1.0, S)
    z = NEAR * (1 - t) + FAR * t
    mid = 0.5 * (z[1:] + z[:-1])
    lo = np.concatenate([z[:1], mid])
    hi = np.concatenate([mid, z[-1:]])
    for (tx, ty) in tiles:
        xs, ys = np.meshgrid(np.arange(tx * tile_w, (tx + 1) * tile_w), np.arange(ty * tile_h, (ty + 1) * tile_h))
        cam = np.stack([(xs + 0.5 - SIDE / 2) / focal, -(ys + 0.5 - SIDE / 2) / focal, -np.ones_like(xs, dtype=np.float64)], -1).reshape(-1, 3)
        d = cam @ rot.T.astype(np.float64)
        o = trans.reshape(1, 3).astype(np.float64)
        u = rng.random((d.shape[0], S))
        zz = lo[None] + (hi - lo)[None] * u
        p = o[:, None, :] + d[:, None, :] * zz[:, :, None]                      # [rays, S, 3]
        inside = np.all((p > -WORLD / 2) & (p < WORLD / 2), axis=-1)
        gi = (p + WORLD / 2) / (WORLD / G) - 0.5
        i0 = np.floor(gi).astype(np.int64)                                       # low corner, may be -1
        for k in ks:
            for s0 in range(0, S - k + 1, k):
                m = inside[:, s0:s0 + k]
                if m.sum() < 16:
                    continue
                c0 = i0[:, s0:s0 + k][m]                                         # [n, 3]
                corners = (c0[:, None, :] + np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)])[None]).reshape(-1, 3)
                corners = corners[np.all((corners >= 0) & (corners < G), axis=1)]
                if corners.shape[0] == 0:
                    continue
                lin = (corners[:, 0] * G + corners[:, 1]) * G + corners[:, 2]
                touched = np.unique(lin).size
                ext = corners.max(0) - corners.min(0) + 1
                box = int(ext.prod())
                out[k]["ratio"].append(touched / box)
                out[k]["box_voxels"].append(box)
                out[k]["touched"].append(touched)
    return out


def summarise(st):
    res = {}
    for k, v in st.items():
        r, b, t = np.array(v["ratio"]), np.array(v["box_voxels"]), np.array(v["touched"])
        res[f"K={k}"] = {
            "windows": int(r.size), "touched_over_box_mean": float(r.mean()), "p10": float(np.percentile(r, 10)), "p50": float(np.percentile(r, 50)),
            "p90": float(np.percentile(r, 90)), "max": float(r.max()), "touched_voxels_mean": float(t.mean()),
            "box_voxels_mean": float(b.mean()), "box_kib_mean_at_112B": float(b.mean() * REC / 1024), "box_kib_p90": float(np.percentile(b, 90) * REC / 1024),
        }
    return res


def main():
    rng = np.random.default_rng(0)
    tiles = [(int(x), int(y)) for x, y in zip(rng.integers(8, 42, 24), rng.integers(15, 85, 24))]
    focal = 1111.11
    report = {"workload": "c3: 256^3 grid, 800x800, 256 jittered spp; CTA tile 16x8 pixels; all in-grid samples (upper bound on use: only ~half contribute)",
              "record_bytes": REC}
    for name, (yaw, pitch) in {"c3_view_yaw30_pitch60": (30.0, 60.0), "oblique_yaw45_pitch45": (45.0, 45.0), "axis_aligned_yaw0_pitch90": (0.0, 90.0)}.items():
        rot, trans = spherical_pose(yaw, pitch, HOTDOG_RADIUS)
        st = tile_stats(rot, trans.reshape(3), focal, tiles, (1, 4, 8, 16), rng)
        report[name] = summarise(st)
    # warp tile (8x4 pixels), the unit the lane-group forward gathers for
    rot, trans = spherical_pose(30.0, 60.0, HOTDOG_RADIUS)
    st = tile_stats(rot, trans.reshape(3), focal, tiles, (1, 4, 8), rng, tile_w=8, tile_h=4)
    report["c3_view_warp_tile_8x4"] = summarise(st)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
