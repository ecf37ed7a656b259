import numpy as np, sys
sys.path.insert(0,'/root/repo')
from radargnn_b200 import synthetic
n=100_000; k=16
fr=synthetic.uniform_square(n, seed=0)
X=fr.X_cc.astype(np.float64)
mn=X.min(0); w=X.max(0)-mn; h=np.sqrt(w[0]*w[1]/(n/4)); gx=int(w[0]//h)+1; gy=int(w[1]//h)+1
cx=np.minimum(((X[:,0]-mn[0])/h).astype(int),gx-1); cy=np.minimum(((X[:,1]-mn[1])/h).astype(int),gy-1)
cell=cy*gx+cx
order=np.lexsort((np.arange(n),cell)); P=X[order]; C=cell[order]
start=np.searchsorted(C,np.arange(gx*gy+1))
def cand_list(q):
    c=C[q]; qy,qx=divmod(c,gx); out=[]; rings=[]
    me=P[q]
    for r in range(0,6):
        x0,x1,y0,y1=qx-r,qx+r,qy-r,qy+r
        spans=[]
        if r==0: spans=[(qy,qx,qx)]
        else:
            spans=[(y0,max(x0,0),min(x1,gx-1)),(y1,max(x0,0),min(x1,gx-1))]
            for yy in range(y0+1,y1):
                for xx in (x0,x1):
                    if 0<=xx<gx: spans.append((yy,xx,xx))
        for (row,xa,xb) in spans:
            if row<0 or row>=gy: continue
            p0,p1=start[row*gx+xa],start[row*gx+xb+1]
            for p in range(p0,p1):
                if p!=q: out.append(p)
        rings.append(len(out))
    d=((P[out]-me)**2).sum(1)
    return np.array(d),rings
def sim(qs,D):
    # returns (cascade executions, baseline executions)
    cands=[cand_list(q) for q in qs]
    L=len(qs)
    lists=[[] for _ in qs]      # current top-k (sorted asc)
    thr=[np.inf]*L; queue=[[] for _ in qs]
    execs=0; base_execs=0
    blist=[[] for _ in qs]; bthr=[np.inf]*L
    maxlen=max(len(c[0]) for c in cands)
    ring_ends=[set(c[1]) for c in cands]
    done=[False]*L
    for i in range(maxlen):
        acc=[False]*L; bacc=[False]*L
        for l in range(L):
            d,rings=cands[l]
            if i>=len(d) or done[l]: continue
            # baseline
            if d[i]<bthr[l] or len(blist[l])<k:
                bacc[l]=True; blist[l].append(d[i]); blist[l].sort(); blist[l]=blist[l][:k]; bthr[l]=blist[l][-1] if len(blist[l])==k else np.inf
            if d[i]<thr[l] or len(lists[l])<k:   # stale threshold
                acc[l]=True
        if any(bacc): base_execs+=1
        if D==0:
            continue
        # queue version
        if any(acc[l] and len(queue[l])==D for l in range(L)):
            mx=max(len(q) for q in queue); execs+=mx
            for l in range(L):
                for v in queue[l]:
                    lists[l].append(v); 
                lists[l].sort(); lists[l]=lists[l][:k]; thr[l]=lists[l][-1] if len(lists[l])==k else np.inf; queue[l]=[]
        for l in range(L):
            if acc[l]: queue[l].append(cands[l][0][i])
        # ring end flush + termination emulate: flush when any lane hits a ring end
        if any((i+1) in ring_ends[l] for l in range(L)):
            mx=max(len(q) for q in queue)
            if mx>0:
                execs+=mx
                for l in range(L):
                    for v in queue[l]: lists[l].append(v)
                    lists[l].sort(); lists[l]=lists[l][:k]; thr[l]=lists[l][-1] if len(lists[l])==k else np.inf; queue[l]=[]
            for l in range(L):
                d,rings=cands[l]
                for ri,re in enumerate(rings):
                    if re==i+1 and len(lists[l])==k:
                        gap=(ri+0.0)*h   # crude: after ring ri searched, unsearched region at least ri*h away (lower bound varies)
                        if lists[l][-1] < gap*gap: done[l]=True
                        if bthr[l] < gap*gap: pass
    return execs, base_execs
rng=np.random.default_rng(0)
tot={D:0 for D in (1,2,3,4)}; base=0
W=12
for wi in range(W):
    q0=int(rng.integers(1000,n-1000))//32*32
    qs=list(range(q0,q0+32))
    for D in (1,2,3,4):
        e,b=sim(qs,D); tot[D]+=e
    base+=b
print("baseline cascade executions per warp %.1f"%(base/W))
for D in tot: print("D=%d: %.1f"%(D,tot[D]/W))
