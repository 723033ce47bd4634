"""Generates the committed golden fixtures in tests/golden/ (run in the build container, where /root/reference is
mounted and oracle/_ref holds the compiled, unmodified reference VFH class).

  vfh_golden.npz    outputs of the REFERENCE move_control::VFH (through oracle/_ref) on seeded pseudo-scan sequences
  himm_golden.npz   outputs of the oracle restatement (pinned by the grid_map known-answer tests) on seeded samples
  ranges_golden.npz Steerer::getRangesFromSubmap restatement on seeded layers/poses

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.util import lidar_samples, random_samples  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def scan_sequence(rng, steps, even_only):
    """Pseudo-scans: a few angular obstacle blobs that drift; sometimes very close (emergency), sometimes none."""
    out = np.full((steps, 361, 2), 5000.0)
    out[:, :, 1] = 0.0
    blobs = [dict(c=rng.uniform(0, 180), w=rng.uniform(3, 50), d=rng.uniform(250, 1600), v=rng.uniform(-3, 3))
             for _ in range(rng.integers(0, 5))]
    for s in range(steps):
        if rng.random() < 0.1:
            blobs.append(dict(c=rng.uniform(0, 180), w=rng.uniform(3, 60), d=rng.uniform(150, 1500), v=rng.uniform(-3, 3)))
        if blobs and rng.random() < 0.08:
            blobs.pop(rng.integers(0, len(blobs)))
        for b in blobs:
            b["c"] += b["v"]
            lo, hi = int(max(0, 2 * (b["c"] - b["w"] / 2))), int(min(360, 2 * (b["c"] + b["w"] / 2)))
            for i in range(lo, hi + 1):
                if even_only and (i & 1):
                    continue
                out[s, i, 0] = min(out[s, i, 0], b["d"] + rng.uniform(-20, 20))
    return out


def vfh_case(seed, steps, even_only=True, **params):
    rng = np.random.default_rng(seed)
    v = O.RefVFH(**params)
    ranges = scan_sequence(rng, steps, even_only)
    speed = rng.integers(-20, v.params["max_speed"] + 1, steps).astype(np.int32)
    gdir = rng.uniform(0, 360, steps).astype(np.float32)
    gdir[rng.random(steps) < 0.5] = 90.0
    gdist = rng.uniform(100, 6000, steps).astype(np.float32)
    tol = np.full(steps, 250.0, np.float32)
    dt = rng.choice([0.2, 0.2, 0.2, 0.05, 0.31, 1.5, 0.0], steps)
    H = v.hist_size
    res = dict(ranges=ranges, speed=speed, gdir=gdir, gdist=gdist, tol=tol, dt=dt,
               out_speed=np.zeros(steps, np.int32), out_turn=np.zeros(steps, np.int32),
               picked=np.zeros(steps, np.float32), origin_hist=np.zeros((steps, H), np.float32),
               hist=np.zeros((steps, H), np.float32), last_binary=np.zeros((steps, H), np.float32),
               blocked=np.zeros(steps, np.float32))
    for s in range(steps):
        cs, ct = v.update(ranges[s], speed[s], gdir[s], gdist[s], tol[s], dt[s])
        st = v.state()
        res["out_speed"][s], res["out_turn"][s], res["picked"][s] = cs, ct, st["picked"]
        res["origin_hist"][s], res["hist"][s], res["last_binary"][s] = st["origin_hist"], st["hist"], st["last_binary"]
        res["blocked"][s] = st["blocked_radius"]
    res["params"] = np.array([float(v.params[k]) for k in O.VFH_PARAM_ORDER])
    return res


VFH_CASES = {
    "w30_default": dict(seed=11, steps=60),
    "w30_allidx": dict(seed=12, steps=40, even_only=False),
    "w33": dict(seed=13, steps=40, window_diameter=33),
    "w30_fast": dict(seed=14, steps=40, max_speed=500, max_turnrate_1ms=20, safety_dist_1ms=120.0),
    "w60_fixed_safety": dict(seed=15, steps=30, window_diameter=60, safety_dist_1ms=10.0, robot_radius=300.0),
}

HIMM_CASES = {
    "c1_random": dict(geom=(10.0, 10.0, 0.05, 0.0, 0.0, 0, 0), seed=21, kind="random", n=900),
    "moving_map": dict(geom=(4.0, 4.0, 0.05, 1.3, -0.7, 17, 63), seed=22, kind="lidar", n=720),
    "odd_size": dict(geom=(6.5, 3.5, 0.05, -2.0, 5.0, 0, 0), seed=23, kind="random", n=500),
}


def main():
    vfh = {}
    for name, kw in VFH_CASES.items():
        for k, v in vfh_case(**kw).items():
            vfh[name + "/" + k] = v
    np.savez_compressed(os.path.join(HERE, "vfh_golden.npz"), **vfh)

    himm = {}
    for name, c in HIMM_CASES.items():
        lx, ly, res, px, py, s0, s1 = c["geom"]
        g = O.make_geom(lx, ly, res, px, py, (s0, s1))
        rng = np.random.default_rng(c["seed"])
        layer = O.new_layer(g)
        batches = []
        for b in range(3):
            s = random_samples(rng, g, c["n"]) if c["kind"] == "random" else \
                lidar_samples(rng, g, (px + 0.2, py - 0.1), c["n"], 0.2, min(lx, ly) * 0.8, clear_frac=0.1)
            O.himm_update(g, layer, s)
            batches.append(s)
        himm[name + "/geom"] = np.array(c["geom"])
        himm[name + "/samples"] = np.concatenate(batches).view(np.uint8)
        himm[name + "/batch"] = np.array([len(b) for b in batches])
        himm[name + "/layer"] = layer
    np.savez_compressed(os.path.join(HERE, "himm_golden.npz"), **himm)

    rg = {}
    rng = np.random.default_rng(31)
    for i, (geom, start) in enumerate([((10.0, 10.0, 0.05, 0.0, 0.0), (0, 0)), ((4.0, 4.0, 0.05, 1.3, -0.7), (17, 63)),
                                       ((6.5, 3.5, 0.05, -2.0, 5.0), (0, 0))]):
        g = O.make_geom(*geom, start)
        layer = O.new_layer(g)
        occ = rng.random(layer.shape)
        layer[occ < 0.10] = 0.0
        layer[(occ >= 0.10) & (occ < 0.13)] = rng.choice([10.0, 30.0, 90.0, 180.0, 3.0, 3.5], ((occ >= 0.10) & (occ < 0.13)).sum())
        poses = np.stack([geom[3] + (rng.random(24) - 0.5) * geom[0] * 1.1, geom[4] + (rng.random(24) - 0.5) * geom[1] * 1.1,
                          rng.uniform(-np.pi, np.pi, 24)], 1)
        ranges = np.stack([O.ranges_from_submap(g, layer, *p) for p in poses])
        rg["%d/geom" % i] = np.array(list(geom) + list(start))
        rg["%d/layer" % i] = layer
        rg["%d/poses" % i] = poses
        rg["%d/ranges" % i] = ranges[:, :, 0]
    np.savez_compressed(os.path.join(HERE, "ranges_golden.npz"), **rg)
    for f in ("vfh_golden.npz", "himm_golden.npz", "ranges_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
