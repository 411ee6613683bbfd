"""Work-count prototype (CPU, numpy): how many box tests / node visits does an N-wide index over the REFERENCE's leaf
boxes need for the 1080p cat frame, against the reference's binary tree? Stats only (plain float32 slab test)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import profiles, scenes

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
sc = scenes.cat_scene("optimized")
p = profiles.params("optimized", W, H, 1, 1)
out = scenes.run_oracle(sc, p, want=("hit_obj", "hit_t"))
print("oracle work", out["work"])
bvh = sc["mesh"][2]
nn = bvh.shape[0]
left = bvh[:, 0].astype(np.int64); right = bvh[:, 1].astype(np.int64)
mn = bvh[:, 2:5]; mx = bvh[:, 5:8]
ts = bvh[:, 8].astype(np.int64); te = bvh[:, 9].astype(np.int64)
is_leaf = left < 0
print("nodes", nn, "leaves", is_leaf.sum())

# rays
jj, ii = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
uc = np.stack([jj - np.float32(W) / 2 + np.float32(0.5), np.float32(H) / 2 - ii - np.float32(0.5), np.full_like(jj, p.z)], -1).reshape(-1, 3)
u = uc / np.linalg.norm(uc, axis=1, keepdims=True)
O = np.tile(np.array([0, 0, 55], np.float32), (u.shape[0], 1))
t = out["hit_t"].reshape(-1)
P = O + t[:, None] * u
L = np.array(profiles.LIGHT[0], np.float32)
su = L - P
D = np.linalg.norm(su, axis=1, keepdims=True)
su = su / D
rays_O = np.concatenate([O, P + 1e-3 * su]).astype(np.float32)
rays_u = np.concatenate([u, su]).astype(np.float32)
tmax = np.concatenate([np.full(len(u), 1e30, np.float32), D[:, 0]])
NR = rays_O.shape[0]
with np.errstate(divide="ignore"):
    inv = (1.0 / rays_u).astype(np.float32)


def slab(ridx, bmn, bmx):
    o = rays_O[ridx]; r = inv[ridx]
    t0 = (bmn - o) * r; t1 = (bmx - o) * r
    lo = np.minimum(t0, t1).max(axis=1); hi = np.maximum(t0, t1).min(axis=1)
    return hi > lo


def count_tree(children, cmn, cmx, cleaf, root_nodes, name):
    """children: (n_nodes, K) int (-1 none; >=0 node index if not leaf; leaf id if cleaf) ; BFS per level."""
    K = children.shape[1]
    ridx = np.arange(NR); rh = slab(ridx, mn[0][None], mx[0][None])
    ridx = ridx[rh]
    fr_r = np.repeat(ridx, len(root_nodes)); fr_n = np.tile(np.asarray(root_nodes), len(ridx))
    visits = 0; tests = 0; leaf_hits = 0; level = 0
    per_ray_visits = np.zeros(NR, np.int64)
    while len(fr_r):
        visits += len(fr_r)
        np.add.at(per_ray_visits, fr_r, 1)
        nr, nnod = [], []
        for k in range(K):
            c = children[fr_n, k]
            ok = c >= 0
            tests += ok.sum()
            rr = fr_r[ok]; nk = fr_n[ok]; cc = c[ok]
            hit = slab(rr, cmn[nk, k], cmx[nk, k])
            lf = cleaf[nk, k]
            leaf_hits += (hit & lf).sum()
            go = hit & ~lf
            nr.append(rr[go]); nnod.append(cc[go])
        fr_r = np.concatenate(nr); fr_n = np.concatenate(nnod)
        level += 1
    print("%-28s levels %2d  node visits %9d  box tests %9d  leaf hits %8d  max visits/ray %d  mean(entering) %.2f" % (
        name, level, visits, tests, leaf_hits, per_ray_visits.max(), visits / max(1, len(ridx))))
    return visits, tests, leaf_hits


# (a) reference binary tree
ch = np.stack([left, right], 1)
inner = np.where(~is_leaf)[0]
cm = np.zeros((nn, 2, 3), np.float32); cM = np.zeros((nn, 2, 3), np.float32); cl = np.zeros((nn, 2), bool)
for k, c in enumerate((left, right)):
    cc = np.where(c >= 0, c, 0)
    cm[:, k] = mn[cc]; cM[:, k] = mx[cc]; cl[:, k] = is_leaf[cc]
ch2 = np.where(is_leaf[:, None], -1, ch)
count_tree(ch2, cm, cM, cl, [0], "reference binary")


# (b) collapse the reference tree to K-wide: repeatedly replace the inner child with the largest surface area by its children
def collapse(K):
    def area(n):
        d = mx[n] - mn[n]
        return d[0] * d[1] + d[1] * d[2] + d[0] * d[2]
    nodes = []  # list of child lists (reference node ids)
    index = {}
    order = [0]
    out_children = []
    while order:
        n = order.pop()
        if n in index:
            continue
        index[n] = len(out_children)
        cs = [left[n], right[n]]
        while len(cs) < K:
            cand = [c for c in cs if not is_leaf[c]]
            if not cand:
                break
            b = max(cand, key=area)
            cs.remove(b); cs += [left[b], right[b]]
        out_children.append(cs)
        for c in cs:
            if not is_leaf[c]:
                order.append(c)
    n_w = len(out_children)
    chw = -np.ones((n_w, K), np.int64); cmn = np.zeros((n_w, K, 3), np.float32); cmx = np.zeros((n_w, K, 3), np.float32); clf = np.zeros((n_w, K), bool)
    for n, cs in zip(list(index.keys()), out_children):
        w = index[n]
        for k, c in enumerate(cs):
            clf[w, k] = is_leaf[c]
            chw[w, k] = c if is_leaf[c] else index[c]
            cmn[w, k] = mn[c]; cmx[w, k] = mx[c]
    return chw, cmn, cmx, clf


if not is_leaf[0]:
    for K in (4, 8):
        chw, cmn, cmx, clf = collapse(K)
        print("collapse K", K, "wide nodes", len(chw), "fill", (chw >= 0).mean())
        count_tree(chw, cmn, cmx, clf, [0], "collapsed ref K=%d" % K)

# (c) own builder over leaf boxes: binary SAH (sweep over centroids), then collapse by area to K-wide
leaf_ids = np.where(is_leaf)[0]
lmn = mn[leaf_ids]; lmx = mx[leaf_ids]
cen = 0.5 * (lmn + lmx)


def sa(a, b):
    d = np.maximum(b - a, 0)
    return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 0] * d[..., 2]


class B:  # binary node of own tree
    pass


def build(ids, mode):
    nd = B(); nd.mn = lmn[ids].min(0); nd.mx = lmx[ids].max(0); nd.ids = ids; nd.l = nd.r = None
    if len(ids) == 1:
        return nd
    best = None
    for ax in range(3):
        o = ids[np.argsort(cen[ids, ax], kind="stable")]
        if mode == "median":
            ext = nd.mx - nd.mn
            if ax != int(np.argmax(ext)):
                continue
            k = len(o) // 2
            best = (0, o[:k], o[k:]); break
        pm = np.minimum.accumulate(lmn[o], 0); pM = np.maximum.accumulate(lmx[o], 0)
        sm_ = np.minimum.accumulate(lmn[o][::-1], 0)[::-1]; sM = np.maximum.accumulate(lmx[o][::-1], 0)[::-1]
        n = len(o)
        cost = sa(pm[:-1], pM[:-1]) * np.arange(1, n) + sa(sm_[1:], sM[1:]) * np.arange(n - 1, 0, -1)
        k = int(np.argmin(cost))
        if best is None or cost[k] < best[0]:
            best = (cost[k], o[:k + 1], o[k + 1:])
    nd.l = build(best[1], mode); nd.r = build(best[2], mode)
    return nd


def widen(root, K):
    wide = []

    def rec(nd):
        me = len(wide); wide.append(None)
        cs = [nd.l, nd.r]
        while len(cs) < K:
            cand = [c for c in cs if c.l is not None]
            if not cand:
                break
            b = max(cand, key=lambda c: sa(c.mn, c.mx))
            cs.remove(b); cs += [b.l, b.r]
        ent = []
        for c in cs:
            if c.l is None:
                ent.append((True, int(c.ids[0]), c.mn, c.mx))
            else:
                ent.append((False, rec(c), c.mn, c.mx))
        wide[me] = ent
        return me
    rec(root)
    n_w = len(wide)
    chw = -np.ones((n_w, K), np.int64); cmn = np.zeros((n_w, K, 3), np.float32); cmx = np.zeros((n_w, K, 3), np.float32); clf = np.zeros((n_w, K), bool)
    for w_, ent in enumerate(wide):
        for k, (lf, c, a, b) in enumerate(ent):
            chw[w_, k] = c; clf[w_, k] = lf; cmn[w_, k] = a; cmx[w_, k] = b
    return chw, cmn, cmx, clf


sys.setrecursionlimit(100000)
if len(leaf_ids) > 1:
    for mode in ("sah", "median"):
        root = build(np.arange(len(leaf_ids)), mode)
        for K in (4, 8, 16):
            chw, cmn, cmx, clf = widen(root, K)
            print(mode, "K", K, "wide nodes", len(chw), "fill %.2f" % (chw >= 0).mean())
            count_tree(chw, cmn, cmx, clf, [0], "own %s K=%d" % (mode, K))
