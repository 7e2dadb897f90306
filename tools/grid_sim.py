"""CPU study for DESIGN.md section 6 item 6: how many pair evaluations an exact uniform-grid nearest-neighbour search needs on
the synthetic clouds (cell block = 27-neighbourhood, exactness check = distance to the block boundary), and how many
queries the first ring leaves unresolved.  Pure numpy; run: python tools/grid_sim.py"""
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
