"""CPU analysis (oracle data, no GPU): how full are the warps of the render kernels?  For a sample of tiles, every
(8x4 patch, list entry) pair in which at least one pixel accumulates the entry ("hit" pair, what the backward sweeps)
and the number of its 32 lanes that do.  DESIGN.md §4 quotes the result.  Usage: python scripts/analysis_lane_util.py lego_1m"""
import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from splatfields_b200 import synth
from tests.helpers import run_oracle
O.build(); O.set_num_threads(8)
name=sys.argv[1] if len(sys.argv)>1 else 'lego_100k'
cfg=synth.CONFIGS[name]
sc=synth.make_scene(cfg['P'],cfg['seed'],scale_mult=cfg['scale_mult'],precomp_rgb=cfg['precomp_rgb']); cam=synth.config_camera(name,0)
H,W=cfg['H'],cfg['W']
f,_=run_oracle(O,sc,cam,H,W,(1,1,1),0 if cfg['precomp_rgb'] else 3,want_margin=False)
gx=(W+15)//16; gy=(H+15)//16
rng=np.random.RandomState(0)
tiles=rng.choice(gx*gy, 300, replace=False)
tot_pairs=0; tot_active=0; hist=np.zeros(33,np.int64); swept_entries=0; pairs_4x4=0; act4=0
for t in tiles:
    r0,r1=f['ranges'][t]; 
    if r1<=r0: continue
    tx,ty=t%gx,t//gx
    ys,xs=np.meshgrid(np.arange(16)+16*ty,np.arange(16)+16*tx,indexing='ij')
    inside=(xs<W)&(ys<H)
    nc=np.zeros((16,16),np.int64); nc[inside]=f['n_contrib'][ys[inside],xs[inside]]
    deepest=int(nc.max()); 
    if deepest==0: continue
    ids=f['point_list'][r0:r0+deepest]
    m=f['means2D'][ids]; co=f['conic_opacity'][ids]
    dx=m[:,0,None,None]-xs[None]; dy=m[:,1,None,None]-ys[None]
    power=-0.5*(co[:,0,None,None]*dx*dx+co[:,2,None,None]*dy*dy)-co[:,1,None,None]*dx*dy
    alpha=np.minimum(0.99,co[:,3,None,None]*np.exp(np.minimum(power,0)))
    pos=np.arange(deepest)[:,None,None]
    contrib=(power<=0)&(alpha>=1/255)&(pos<nc[None])    # entries each pixel actually accumulates (before its last contributor)
    swept_entries+=deepest
    # 8x4 patches: warp w -> rows (w>>1)*4.., cols (w&1)*8..
    c=contrib.reshape(deepest,4,4,2,8)   # [e, wy, ly, wx, lx]
    per=c.sum(axis=(2,4))                # [e, wy, wx] active lanes
    hit=per>0
    tot_pairs+=hit.sum(); tot_active+=per[hit].sum()
    hist+=np.bincount(per[hit].reshape(-1),minlength=33)
    c4=contrib.reshape(deepest,4,4,4,4).sum(axis=(2,4)); h4=c4>0; pairs_4x4+=h4.sum(); act4+=c4[h4].sum()
print(name,'tiles sampled',len(tiles),'entries swept/tile',swept_entries/len(tiles))
print('8x4 patches: hit pairs/tile',tot_pairs/len(tiles),' mean active lanes of 32:',tot_active/tot_pairs)
print('lane histogram (1..32) cumulative %:',np.round(100*np.cumsum(hist[1:])/hist[1:].sum(),1)[[0,1,3,7,11,15,23,31]])
print('4x4 patches: hit pairs/tile',pairs_4x4/len(tiles),' mean active of 16:',act4/pairs_4x4)
