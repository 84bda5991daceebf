import sys; sys.path.insert(0,'.')
import torch, numpy as np
from dgod_b200 import synth
B,H,W,per=8,608,1024,512
tot_rows=[];tot_span=[];lv=[]
for i in range(B):
    b=synth.random_boxes(per,H,W,synth.gen(10+i)).numpy().astype(np.float64)
    area=(b[:,2]-b[:,0])*(b[:,3]-b[:,1])
    k=np.clip(np.floor(4+np.log2(np.sqrt(area)/224)+1e-6),2,5).astype(int)
    s=1.0/2.0**k
    for j in range(per):
        x1,y1,x2,y2=b[j]*s[j]
        rw=max(x2-x1,1);rh=max(y2-y1,1)
        Hl=-(-H//2**k[j]);Wl=-(-W//2**k[j])
        def axis(start,size,n):
            bins=size/7
            cs=[start+p*bins+(q+.5)*bins/2 for p in range(7) for q in range(2)]
            live=set()
            for c in cs:
                if c<-1 or c>n: continue
                c=max(c,0);lo=int(c)
                if lo>=n-1: lo=n-1;hi=n-1
                else: hi=lo+1
                live.add(lo);
                if c-lo>0: live.add(hi)
            return live
        r=axis(y1,rh,Hl);c=axis(x1,rw,Wl)
        tot_rows.append(len(r));tot_span.append(max(c)-min(c)+1 if c else 0);lv.append(k[j])
tr=np.array(tot_rows);ts=np.array(tot_span);lv=np.array(lv)
print("rows mean",tr.mean(),"span mean",ts.mean(),"max",tr.max(),ts.max())
print("pixels per roi mean",(tr*ts).mean(), "span>28:",(ts>28).mean())
print("levels",np.bincount(lv))
print("sum rows*span*1KB MB", (tr*ts).sum()*1024/1e6)
for l in range(2,6): print(l,tr[lv==l].mean(),ts[lv==l].mean())
for t in (28, 32, 36, 40, 44, 48): print("span >", t, (ts > t).mean(), " share of px:", (tr*ts)[ts > t].sum() / (tr*ts).sum())
