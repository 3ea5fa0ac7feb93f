#!/usr/bin/env bash
# K1 alone at capped occupancy: extra dynamic smem per CTA limits the CTAs per SM (228 KB / (8 KB + extra) at R = 15).
for R in 15 19 25; do
for X in 0 11000 20000 30000 38000; do
  DH_RRC_R=$R DH_RRC_EXTRA_SMEM=$X python - <<PY
import os,sys,torch
sys.path.insert(0,'.')
import digiham_b200 as dh
C,L=4096,48000
x=torch.rand((C,L),device='cuda')*2-1
bank=dh.RrcBank(C)
out=torch.empty((C,L),device='cuda')
for _ in range(3): bank.process(x,out=out,n=L)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): bank.process(x,out=out,n=L)
e1.record(); torch.cuda.synchronize()
R=int(os.environ['DH_RRC_R']); X=int(os.environ['DH_RRC_EXTRA_SMEM'])
smem=(80+128*R)*4+X
print("R=%d extra smem %5d B -> %2d CTAs/SM by smem: %.4f ms" % (R, X, min(227*1024//(smem+1024), 16), e0.elapsed_time(e1)/20))
PY
done; done
