"""CPU study for DESIGN.md section 6 item 6: how many pair evaluations an exact uniform-grid nearest-neighbour search needs on
the synthetic clouds (cell block = 27-neighbourhood, exactness check = distance to the block boundary), and how many
queries the first ring leaves unresolved.  Pure numpy; run: python tools/grid_sim.py [--exact]
(--exact: per-query ring search in fp32 with the tie rule, checked bit for bit against brute force)."""
import numpy as np, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import synthetic
def sim(a, c, h):
    # a: queries (n,3), c: refs (m,3); uniform grid over the joint bounding box
    lo = np.minimum(a.min(0), c.min(0)) - 1e-6
    ca = np.floor((a - lo) / h).astype(int); cc = np.floor((c - lo) / h).astype(int)
    dims = np.maximum(ca.max(0), cc.max(0)) + 1
    def key(x): return (x[:,0]*dims[1] + x[:,1])*dims[2] + x[:,2]
    cnt_c = np.zeros(dims, int); np.add.at(cnt_c, tuple(cc.T), 1)
    cnt_a = np.zeros(dims, int); np.add.at(cnt_a, tuple(ca.T), 1)
    # 27-neighbourhood counts by box filter
    pad = np.pad(cnt_c, 1)
    K = np.zeros(dims, int)
    for dx in range(3):
        for dy in range(3):
            for dz in range(3):
                K += pad[dx:dx+dims[0], dy:dy+dims[1], dz:dz+dims[2]]
    pairs = int((cnt_a * K).sum())
    # exact check per query
    d2 = ((a[:,None,:]-c[None,:,:])**2).sum(-1)
    nn = d2.min(1)
    near = (np.abs(ca[:,None,:]-cc[None,:,:]).max(-1) <= 1)
    nn27 = np.where(near, d2, np.inf).min(1)
    frac = (a - lo)/h - ca
    margin = h*(1+np.minimum(frac, 1-frac).min(1))
    resolved = nn27 <= margin**2
    assert (nn27[resolved] == nn[resolved]).all()
    return pairs, int((~resolved).sum()), int((cnt_a>0).sum()), int(cnt_a.max()), int(K.max())
for name, mk in (("S1", synthetic.s1_near), ("S2", synthetic.s2_far)):
    A, C = mk(4, 2048)
    A, C = A.numpy().astype(np.float64), C.numpy().astype(np.float64)
    for h in (0.04, 0.06, 0.08, 0.1, 0.15):
        tot=[sim(A[b], C[b], h) for b in range(4)]
        p=np.mean([t[0] for t in tot]); u=np.mean([t[1] for t in tot])
        print(f"{name} h={h}: pairs/(n*m)={p/2048/2048:.3f}  unresolved queries={u:.0f}/2048 ({u*2048/2048/2048:.3f} of n*m if brute-forced)  nonempty cells={np.mean([t[2] for t in tot]):.0f} max q/cell={max(t[3] for t in tot)} max K={max(t[4] for t in tot)}")


# ---- exactness of the pruned search incl. the tie rule (fp32 distances, lowest index among equal distances) --------
def d2_f32(q, r):
    """fp32 squared distance in the kernels' order fma(dz,dz,fma(dx,dx,dy*dy)), d* = r - q (fma emulated through float64)."""
    q, r = q.astype(np.float32), r.astype(np.float32)
    dx, dy, dz = (r[..., 0] - q[..., 0]).astype(np.float32), (r[..., 1] - q[..., 1]).astype(np.float32), \
        (r[..., 2] - q[..., 2]).astype(np.float32)
    t = (dy * dy).astype(np.float32)
    t = (dx.astype(np.float64) * dx + t).astype(np.float32)
    return (dz.astype(np.float64) * dz + t).astype(np.float32)


def grid_nn(a, c, h):
    """Exact nearest neighbour of every a[i] in c by ring search on a uniform grid; returns (dist, idx, pair evaluations)."""
    lo = np.minimum(a.min(0), c.min(0)).astype(np.float64) - 1e-6
    ca = np.floor((a - lo) / h).astype(int)
    cc = np.floor((c - lo) / h).astype(int)
    frac = (a - lo) / h - ca
    inner = np.minimum(frac, 1 - frac).min(1)  # distance (in cells) to the nearest face of the own cell
    n = len(a)
    dist, idx, pairs = np.empty(n, np.float32), np.empty(n, np.int64), 0
    rmax = int(np.abs(ca[:, None, :] - cc[None, :, :]).max()) + 1
    for i in range(n):
        cheb = np.abs(cc - ca[i]).max(1)
        best, bi, ring = np.float32(np.inf), -1, 0
        while True:
            ring += 1
            cand = np.flatnonzero(cheb <= 1) if ring == 1 else np.flatnonzero(cheb == ring)  # 27-cell block, then shells
            pairs += len(cand)
            if len(cand):
                d = d2_f32(a[i][None, :], c[cand])
                j = int(np.lexsort((cand, d))[0])  # smallest distance, then lowest index
                if d[j] < best or (d[j] == best and cand[j] < bi):
                    best, bi = d[j], int(cand[j])
            # every point outside the (2 ring + 1)^3 block is at least h (ring + inner) away; keep a relative margin far
            # above the rounding error of the fp32 distance (3 ulp) and stop only on STRICT inequality (tie rule)
            bound = (h * (ring + inner[i])) ** 2 * (1 - 2.0 ** -18)
            if float(best) < bound or ring > rmax:
                break
        dist[i], idx[i] = best, bi
    return dist, idx, pairs


if __name__ == "__main__" and "--exact" in sys.argv:
    for name, (A, C), h in (("S1", synthetic.s1_near(1, 1024), 0.08), ("S2", synthetic.s2_far(1, 700, 900), 0.12),
                            ("S3 ties", synthetic.s3_ties(1, 1024, pool=256), 0.1)):
        a, c = A[0].numpy(), C[0].numpy()
        for q, r, tag in ((a, c, "1->2"), (c, a, "2->1")):
            dist, idx, pairs = grid_nn(q, r, h)
            dense = d2_f32(q[:, None, :], r[None, :, :])
            want_i = np.array([int(np.lexsort((np.arange(len(r)), dense[i]))[0]) for i in range(len(q))])
            assert np.array_equal(idx, want_i) and np.array_equal(dist, dense[np.arange(len(q)), want_i]), (name, tag)
            print(f"{name} {tag}: bit-identical to brute force (indices incl. ties), {pairs / dense.size:.3f} of the pair evaluations")
