"""CPU simulation behind the culled EMD sweeps (csrc/approxmatch.cu): which fraction of the partners lies within the
underflow radius of level j of a block's bounding box, for Morton-ordered and for median-split (k-d) blocks of 32 / 64
points, on the S1 and S2 clouds; and the fraction of PAIRS inside the radius (what per-pair skipping could reach)."""
import numpy as np, torch, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import synthetic
L2E=1.4426950408889634
def morton(p, bits=10):
    lo=p.min(0); hi=p.max(0)
    q=((p-lo)/(hi-lo+1e-12)*(2**bits-1)).astype(np.uint64)
    code=np.zeros(len(p),dtype=np.uint64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:,a]>>np.uint64(b))&np.uint64(1))<<np.uint64(3*b+a)
    return np.argsort(code,kind='stable')
def frac(own, part, R2, group=64, order='morton'):
    perm = morton(own) if order=='morton' else np.arange(len(own))
    o=own[perm]
    fr=[]
    for g in range(0,len(o),group):
        blk=o[g:g+group]; lo=blk.min(0); hi=blk.max(0)
        d=np.maximum(np.maximum(lo-part,0),part-hi)
        d2=(d*d).sum(1)
        fr.append((d2<R2).mean())
    return np.array(fr)
for name,gen in [('S1',lambda: synthetic.s1_near(4,2048)),('S2',lambda: synthetic.s2_far(4,2048))]:
    a,b=gen(); a=a.numpy().astype(np.float64); b=b.numpy().astype(np.float64)
    for j in [7,6,5,4,3]:
        lc=4.0**j*L2E; R2=130/lc
        for group in (64,32):
            f1=np.concatenate([frac(a[i],b[i],R2,group) for i in range(4)])
            f2=np.concatenate([frac(b[i],a[i],R2,group) for i in range(4)])
            print(name,'j',j,'R',round(R2**.5,3),'group',group,'own=1 mean %.3f max %.3f | own=2 mean %.3f max %.3f'%(f1.mean(),f1.max(),f2.mean(),f2.max()))
print('--- kd grouping')
def kd_groups(p, idx, group):
    if len(idx)<=group: return [idx]
    ext=p[idx].max(0)-p[idx].min(0); a=int(np.argmax(ext))
    o=idx[np.argsort(p[idx,a],kind='stable')]; h=len(o)//2
    return kd_groups(p,o[:h],group)+kd_groups(p,o[h:],group)
def frac_kd(own, part, R2, group):
    fr=[]
    for g in kd_groups(own,np.arange(len(own)),group):
        blk=own[g]; lo=blk.min(0); hi=blk.max(0)
        d=np.maximum(np.maximum(lo-part,0),part-hi)
        fr.append(((d*d).sum(1)<R2).mean())
    return np.array(fr)
for name,gen in [('S1',lambda: synthetic.s1_near(4,2048)),('S2',lambda: synthetic.s2_far(4,2048))]:
    a,b=gen(); a=a.numpy().astype(np.float64); b=b.numpy().astype(np.float64)
    for j in [7,6,5,4]:
        lc=4.0**j*L2E; R2=130/lc
        pairfrac=np.mean([(((a[i][:,None,:]-b[i][None,:,:])**2).sum(-1)<R2).mean() for i in range(4)])
        for group in (64,32):
            f1=np.concatenate([frac_kd(a[i],b[i],R2,group) for i in range(4)])
            print(name,'j',j,'group',group,'mean %.3f max %.3f p90 %.3f pairfrac %.4f'%(f1.mean(),f1.max(),np.quantile(f1,.9),pairfrac))
