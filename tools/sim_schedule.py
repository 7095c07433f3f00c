"""Experiments only: list-scheduling simulation of one bench step from a per-warp timeline (tools/timeline.py, FRIQ_GROUPS=1):
what would other group layouts (index groups, groups stratified by rule count) do to the step time?  Answer for the table-mode
kernel: nothing - the step is bounded by the two launches of the tile with the largest rule bases (critical path)."""
import numpy as np, heapq, sys
d=np.load('/root/repo/gpurun_out/timeline_g1r125.npy')
t0,t1,smid,wid,rmax=d[:,0],d[:,1],d[:,2],d[:,3],d[:,4]
# two launches per warp id
L=[{},{}]
for i in np.argsort(t0):
    w=int(wid[i]); k=0 if w not in L[0] else 1
    L[k][w]=(t1[i]-t0[i], rmax[i])
nw=len(L[0]); ncta=nw//4
def cta_jobs(l):
    dur=np.array([max(L[l][c*4+j][0] for j in range(4)) for c in range(ncta)])/1e3   # us
    r=np.array([max(L[l][c*4+j][1] for j in range(4)) for c in range(ncta)])
    return dur,r
d1,r1=cta_jobs(0); d2,r2=cta_jobs(1)
print('ctas',ncta,'launch1 max/mean',d1.max(),d1.mean(),'launch2',d2.max(),d2.mean(), 'sum work (slot-us)', d1.sum()+d2.sum(), 'ideal ms', (d1.sum()+d2.sum())/296/1e3)
SLOTS=296; RET=80.0
def simulate(groups):
    """groups: list of lists of (dur1,dur2) arrays in dispatch order; each group: launch1 -> RET -> launch2.  CTAs dispatched
    greedily to free slots; kernels ready in (group, launch) order, earlier-ready first."""
    # event simulation
    free=[0.0]*SLOTS; heapq.heapify(free)
    # per group state
    pend=[]  # (ready_time, group, launch, idx)
    done_time={}
    queue=[]  # ready kernels: list of [ready_time, g, l, next_idx]
    for g,(a,b) in enumerate(groups): queue.append([RET,g,0,0])
    finish={}
    remaining={}
    for g,(a,b) in enumerate(groups): remaining[(g,0)]=len(a); remaining[(g,1)]=len(b); finish[(g,0)]=0; finish[(g,1)]=0
    t_end=0
    active=queue
    while active:
        # pick the slot that frees first
        t=heapq.heappop(free)
        # choose kernel with ready_time<=t having CTAs left, earliest ready first; else advance t
        cand=[k for k in active if k[0]<=t]
        if not cand:
            tn=min(k[0] for k in active); heapq.heappush(free,max(t,tn)); continue
        k=min(cand,key=lambda x:(x[0],x[1]))
        g,l,i=k[1],k[2],k[3]
        dur=groups[g][l][i]
        end=t+dur; heapq.heappush(free,end)
        finish[(g,l)]=max(finish[(g,l)],end); t_end=max(t_end,end)
        k[3]+=1
        if k[3]>=len(groups[g][l]):
            active.remove(k)
            if l==0: active.append([None,g,1,0]); 
        # resolve ready time of launch 2 once all CTAs of launch 1 are dispatched -> its finish known
        for kk in active:
            if kk[0] is None: kk[0]=finish[(kk[1],0)]+RET
    return t_end
# S0: current: 2 index groups, each a mix: emulate by interleaving sorted CTAs alternately
order=np.argsort(-r1)  # heavy first
def mk(idx): return (d1[idx], d2[idx])
print('g1 (one group)        ', simulate([mk(order)])/1e3)
print('2 index groups (mixed)', simulate([mk(order[0::2]),mk(order[1::2])])/1e3)
print('4 index groups (mixed)', simulate([mk(order[j::4]) for j in range(4)])/1e3)
h=len(order)//2
print('2 stratified groups   ', simulate([mk(order[:h]),mk(order[h:])])/1e3)
q=len(order)//4
print('4 stratified groups   ', simulate([mk(order[j*q:(j+1)*q]) for j in range(4)])/1e3)
print('8 stratified groups   ', simulate([mk(order[j*q//2:(j+1)*q//2]) for j in range(8)])/1e3)
