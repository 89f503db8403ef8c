import sys, numpy as np, warnings
warnings.filterwarnings("ignore")
tr=np.load(sys.argv[1]).astype(np.float64)
t0=tr[tr>0].min()
tr=np.where(tr>0,(tr-t0)/1e3,np.nan)
n,T,K=tr.shape
names=["issue","simstart","arrive","flag","posd","Astart","Aend","flushed"]
for t in [0,1,2,3,8,16,32,48,62,63]:
    print(f"t={t:2d} "+" ".join(f"{nm}={np.nanmedian(tr[:,t,k]):6.1f}" for k,nm in enumerate(names)))
x=tr[:,8:60,:]
def d(a,b): return x[:,:,b]-x[:,:,a]
for nm,(a,b) in {"issue->arrive":(0,2),"simstart->arrive":(1,2),"arrive->flag":(2,3),"flag->gap":(3,4),"flag->mstart":(3,5),"mstart->mdone":(5,6),"posd->flushed":(4,7),"Aend->flushed":(6,7)}.items():
    print(f"{nm:18s} med {np.nanmedian(d(a,b)):6.2f} mean {np.nanmean(d(a,b)):6.2f} p95 {np.nanpercentile(d(a,b),95):6.2f}")
for k,nm in ((6,"merge"),(3,"flag"),(0,"issue")):
    per=tr[:,9:60,k]-tr[:,8:59,k]
    print(f"{nm} period med {np.nanmedian(per):.2f} mean {np.nanmean(per):.2f}")
print("span", np.nanmax(tr))
