"""Prototype (CPU, numpy) of anchored-ray binning: for rays whose line passes through a fixed anchor point (camera
rays; shadow rays, which end at the light), the direction space around the anchor (3 faces x R x R cells, d and -d
folded together) lists the reference leaves whose box a ray of that cell can hit. Checks the superset property against
a float32 slab test on the real rays of config 2 and prints the candidate statistics."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import profiles, scenes

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
R = int(sys.argv[3]) if len(sys.argv) > 3 else 256
sc = scenes.cat_scene("optimized")
p = profiles.params("optimized", W, H, 1, 1)
out = scenes.run_oracle(sc, p, want=("hit_obj", "hit_t"))
bvh = sc["mesh"][2]
is_leaf = bvh[:, 0] < 0
lmn = bvh[is_leaf, 2:5].astype(np.float64); lmx = bvh[is_leaf, 5:8].astype(np.float64)
nl = len(lmn)
print("leaves", nl)

jj, ii = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
uc = np.stack([jj - np.float32(W) / 2 + np.float32(0.5), np.float32(H) / 2 - ii - np.float32(0.5), np.full_like(jj, p.z)], -1).reshape(-1, 3)
u = (uc / np.linalg.norm(uc, axis=1, keepdims=True)).astype(np.float32)
cam = np.array([0, 0, 55], np.float32)
O = np.tile(cam, (u.shape[0], 1))
t = out["hit_t"].reshape(-1)
P = O + t[:, None] * u
L = np.array(profiles.LIGHT[0], np.float32)
toL = L - P
su = (toL / np.linalg.norm(toL, axis=1, keepdims=True)).astype(np.float32)


def slab_hits(Or, ur):
    """(n_rays, n_leaves) bool via float32 slab test (chunked)"""
    res = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for s in range(0, len(Or), 20000):
            o = Or[s:s + 20000, None, :]; d = ur[s:s + 20000, None, :]
            t0 = (lmn[None].astype(np.float32) - o) / d; t1 = (lmx[None].astype(np.float32) - o) / d
            lo = np.minimum(t0, t1).max(-1); hi = np.maximum(t0, t1).min(-1)
            res.append(np.packbits(hi > lo, axis=1))
    return np.concatenate(res)


def build_bins(A, R, eps):
    """lists[face][cy][cx] -> python list of leaf ids; everywhere: list"""
    lists = [[[[] for _ in range(R)] for _ in range(R)] for _ in range(3)]
    everywhere = []
    for l in range(nl):
        v0 = lmn[l] - eps - A; v1 = lmx[l] + eps - A
        for k in range(3):
            a, b = (k + 1) % 3, (k + 2) % 3
            mina = 0.0 if v0[a] <= 0 <= v1[a] else min(abs(v0[a]), abs(v1[a]))
            minb = 0.0 if v0[b] <= 0 <= v1[b] else min(abs(v0[b]), abs(v1[b]))
            m = max(mina, minb)
            parts = []
            if m == 0.0 and v0[k] <= 0 <= v1[k]:
                everywhere.append(l); break
            if v1[k] > 0 and v1[k] >= m:
                parts.append((max(v0[k], m, 1e-30), v1[k]))
            if v0[k] < 0 and -v0[k] >= m:
                parts.append((v0[k], min(v1[k], -m, -1e-30)))
            for (k0, k1) in parts:
                ra = [x / y for x in (v0[a], v1[a]) for y in (k0, k1)]
                rb = [x / y for x in (v0[b], v1[b]) for y in (k0, k1)]
                a0, a1 = max(min(ra), -1.0), min(max(ra), 1.0)
                b0, b1 = max(min(rb), -1.0), min(max(rb), 1.0)
                if a0 > a1 or b0 > b1:
                    continue
                ca0 = max(int(np.floor((a0 + 1) * 0.5 * R)) - 1, 0); ca1 = min(int(np.floor((a1 + 1) * 0.5 * R)) + 1, R - 1)
                cb0 = max(int(np.floor((b0 + 1) * 0.5 * R)) - 1, 0); cb1 = min(int(np.floor((b1 + 1) * 0.5 * R)) + 1, R - 1)
                for cb in range(cb0, cb1 + 1):
                    row = lists[k][cb]
                    for ca in range(ca0, ca1 + 1):
                        if not row[ca] or row[ca][-1] != l:
                            row[ca].append(l)
    return lists, everywhere


def ray_cells(d, R):
    ad = np.abs(d)
    k = np.argmax(ad, axis=1)
    idx = np.arange(len(d))
    dk = d[idx, k]
    ra = d[idx, (k + 1) % 3] / dk; rb = d[idx, (k + 2) % 3] / dk
    ca = np.clip(np.floor((ra + 1) * 0.5 * R).astype(np.int64), 0, R - 1)
    cb = np.clip(np.floor((rb + 1) * 0.5 * R).astype(np.int64), 0, R - 1)
    return k, cb, ca


def check(name, A, Or, ur, d, R):
    scale = max(np.abs(lmn).max(), np.abs(lmx).max()) + np.abs(A).max()
    eps = scale * 2.0 ** -12
    lists, everywhere = build_bins(A.astype(np.float64), R, eps)
    sizes = np.array([[[len(c) for c in row] for row in face] for face in lists])
    print("%s: R %d eps %.4g everywhere %d  total entries %d  nonempty cells %d  max list %d" % (name, R, eps, len(everywhere), sizes.sum(), (sizes > 0).sum(), sizes.max()))
    k, cb, ca = ray_cells(d, R)
    hits = slab_hits(Or, ur)
    nh = np.unpackbits(hits, axis=1)[:, :nl]
    cand = sizes[k, cb, ca] + len(everywhere)
    true_hits = nh.sum(1)
    print("   rays %d  with candidates %d  mean candidates (all rays) %.2f  (rays with any) %.2f  max %d ; true leaf hits total %d  (per ray with any: %.2f)" % (
        len(d), (cand > 0).sum(), cand.mean(), cand[cand > 0].mean(), cand.max(), true_hits.sum(), true_hits[true_hits > 0].mean()))
    # superset check
    bad = 0
    ev = set(everywhere)
    rows = np.nonzero(true_hits)[0]
    for r in rows:
        cl = set(lists[k[r]][cb[r]][ca[r]]) | ev
        hl = np.nonzero(nh[r])[0]
        for l in hl:
            if l not in cl:
                bad += 1
    print("   superset violations:", bad)


check("camera", cam, O, u, u, R)
check("light", L, (P + np.float32(1e-4) * su).astype(np.float32), su, (P - L).astype(np.float32), R)
