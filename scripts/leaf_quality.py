import sys, importlib, numpy as np
sys.path.insert(0,'/root/repo')
synth = importlib.import_module("spatial-temporal-lidar-camera-calibration_b200.synth")
pack,xgt,_ = synth.generate(n_kf=1)
P = pack.scan_xyz.astype(np.float64); N=len(P)
def spread(v):
    x=v.astype(np.uint64)&0xffff
    x=(x|(x<<16))&0x0000ff0000ff; x=(x|(x<<8))&0x00f00f00f00f; x=(x|(x<<4))&0x0c30c30c30c3; x=(x|(x<<2))&0x249249249249
    return x
lo=P.min(0); ext=(P.max(0)-lo).max(); q=np.minimum(((P-lo)*(65535/ext)).astype(np.int64),65535)
mort = spread(q[:,0])|(spread(q[:,1])<<np.uint64(1))|(spread(q[:,2])<<np.uint64(2))
order_m = np.argsort(mort, kind='stable')
def kd_order(idx, leaf=32):
    # recursive median split on longest axis until <= leaf, keeping leaves exactly `leaf`-aligned
    out=[]
    def rec(ix):
        if len(ix)<=leaf: out.append(ix); return
        pts=P[ix]; ax=np.argmax(pts.max(0)-pts.min(0))
        half=((len(ix)//leaf+1)//2)*leaf if len(ix)%leaf==0 or True else len(ix)//2
        half=min(max(half,leaf),len(ix)-1)
        o=np.argsort(pts[:,ax],kind='stable'); rec(ix[o[:half]]); rec(ix[o[half:]])
    rec(idx); return np.concatenate(out)
def leaves(order):
    n=len(order)//32*32; Q=P[order[:n]].reshape(-1,32,3); return Q.min(1),Q.max(1)
def count_boxes(lo_,hi_,qs,dn):
    cnt=[]
    for qq,d in zip(qs,dn):
        dd=np.maximum(np.maximum(lo_-qq,qq-hi_),0); lb=(dd**2).sum(1); cnt.append((lb<=d).sum())
    return np.mean(cnt), np.percentile(cnt,[50,90,99])
rng=np.random.default_rng(0)
sel=rng.choice(N,400,replace=False); qs=P[sel]+rng.normal(0,0.03,(400,3))
from scipy.spatial import cKDTree
T=cKDTree(P); dn,_=T.query(qs); dn=dn**2
d30,_=T.query(P[sel],k=30); r30=np.minimum(d30[:,-1]**2,0.36)
for name,order in (("morton",order_m),("kd-full",kd_order(np.arange(N))),("kd-in-1024", np.concatenate([kd_order(order_m[i:i+1024]) for i in range(0,N,1024)]))):
    l,h=leaves(order)
    print(name,"1-NN boxes", count_boxes(l,h,qs,dn), " kNN30 boxes", count_boxes(l,h,P[sel],r30), "mean leaf diag", np.linalg.norm(h-l,axis=1).mean())
