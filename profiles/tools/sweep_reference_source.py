"""Parameter sweep of the restated extractFeatures (oracle.extract) against the reference's own CirclesEventFrame.cpp compiled in place
(oracle/_ref/libref_functor.so): 2 sensors x 2 streams (5 % noise; 20 % noise + 5 % polarity flips) x eps {2,3,4,4.5,6} x minPts {2,3,5} x
(clusterMinSample, knn_num, fitCircle) in 5 combinations.  Last run: 1350 windows, 581 with the grid found, 0 mismatches (candidate lists
equal, features bit-identical).  Needs /root/reference at build time only:  python profiles/tools/sweep_reference_source.py"""
import numpy as np, sys
import os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import oracle
from eventcalib_b200 import synth
from test_circles_grid import _lib as grid_lib, _order as grid_order
glib=grid_lib()
tot=found=bad=0
rng=np.random.default_rng(1)
cases=[(346,260,2e6,1.5e-3),(640,480,6e6,2e-3)]
for (W,H,rate,wl) in cases:
    for seed in (3,1003):
        dur=0.02
        ev=synth.make_stream(int(rate*dur),W,H,t0=5.0,duration=dur,seed=seed,noise_frac=0.05 if seed==3 else 0.2,flip_frac=0.0 if seed==3 else 0.05)
        t,x,y,p=ev['t'],ev['x'],ev['y'],ev['p']
        for eps in (2.0,3.0,4.0,4.5,6.0):
            for minS in (2,3,5):
                for (cmin,knn,fit) in ((5,3,1),(5,1,1),(8,5,1),(5,3,0),(3,3,0)):
                    for w in synth.tiling_windows(5.0,5.0+dur,wl)[::3]:
                        a,b=float(w[0]),float(w[1])
                        m=(t>=a-1e-3)&(t<=b+1e-3)
                        r1=oracle.ref_extract(t[m],x[m],y[m],p[m],a,b,W,H,fit,eps=eps,minS=minS,clusterMin=cmin,knn_num=knn)
                        P0,N0,_,_=oracle.event_frame(t,x,y,p,a,b)
                        r0=oracle.extract(P0,N0,eps=eps,minS=minS,clusterMin=cmin,knn_num=knn,fitCircle=fit,Rthr=r1['rthr'])
                        tot+=1
                        if r1['cand_f32'] is None:
                            if r0['enough']: bad+=1; print("enough mismatch",W,seed,eps,minS,cmin,knn,fit,a)
                            continue
                        c0=r0['cand']
                        if len(c0)!=len(r1['cand_f32']) or not np.array_equal(c0[:,2:4].astype(np.float32),r1['cand_f32']):
                            bad+=1; print("cand mismatch",W,seed,eps,minS,cmin,knn,fit,a,len(c0),len(r1['cand_f32'])); continue
                        ok,order=grid_order(glib,c0[:,2:4].astype(np.float32).astype(np.float64))
                        if ok!=r1['found']: bad+=1; print("found mismatch"); continue
                        if ok:
                            found+=1
                            if not np.array_equal(c0[order][:,2:5],r1['features']): bad+=1; print("feature mismatch",W,seed,eps,minS,cmin,knn,fit,a,np.abs(c0[order][:,2:5]-r1['features']).max())
print("windows",tot,"found",found,"mismatches",bad)
