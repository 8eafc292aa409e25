#!/usr/bin/env python
"""Host-side model of phase 2 of sketch_filter_kernel on the benchmark workload (no GPU): for every tile,
the hits each lane owns (a hit = a k-mer start whose leading b bits equal those of some rand[l]) - the
warp runs phase 2 for max-over-lanes iterations today and for ceil(mean) with the balanced variant
(NSMH_SKETCH_BALANCED=1).  Tile geometry is simplified (tiles start at the read's first base).

    python tools/models/phase2_lane_balance.py [lambda_log2 ...]      (default 2 3)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nanospring_b200 as ns
K,N=23,60
rnd=ns.rand_from_seed(20261017,N)
lengths=ns.synth_lengths(3000,10000,seed=1000)
rd=ns.synth_reads_host(lengths, ns.synth_params(genome_len=50_000_000))
code=((rd.bases&2)|((rd.bases&4)>>2)).astype(np.uint64)
mask=(1<<(2*K))-1
tile_words=640
LAMS=[int(x) for x in sys.argv[1:]] or [2, 3]
for LAM in LAMS:
  tot_max=0; tot_mean=0.0; tot_hits=0; tiles=0; tot_pos=0; missing=0.0
  for i in range(rd.numReads):
    a,b=int(rd.offsets[i]),int(rd.offsets[i+1]); L=b-a
    if L<K: continue
    nk=L-K+1
    bbits=min(max(int(np.floor(np.log2(nk)))-LAM,0),11,2*K)
    c=code[a:b]
    # top bbits of kmer at p = first ceil(bbits/2) bases... compute window value of bbits bits
    nb=(bbits+1)//2
    if bbits==0:
        hit=np.ones(nk,dtype=bool)
    else:
        w=np.zeros(nk,dtype=np.uint64)
        for j in range(nb):
            w=(w<<np.uint64(2))|c[j:j+nk]
        if bbits%2: w>>=np.uint64(1)
        targets=set(int((int(r)&mask)>>(2*K-bbits)) for r in rnd)
        hit=np.isin(w,np.array(sorted(targets),dtype=np.uint64))
        present=set(np.unique(w).tolist())
        missing+=sum(1 for r in rnd if int((int(r)&mask)>>(2*K-bbits)) not in present)
    # tiles: global word alignment ignored (start at read start, word 0): approximate geometry
    pos=np.flatnonzero(hit)
    ntile=(nk+tile_words*16-1)//(tile_words*16)
    for t in range(ntile):
        p0=t*tile_words*16; p1=min(nk,p0+tile_words*16)
        pp=pos[(pos>=p0)&(pos<p1)]-p0
        word=pp//16
        lane=(word%64)//2
        cnt=np.bincount(lane,minlength=32)
        tot_max+=cnt.max(); tot_mean+=cnt.sum()/32; tot_hits+=cnt.sum(); tiles+=1; tot_pos+=p1-p0
  print(f"lambda_log2={LAM}: reads {rd.numReads} tiles {tiles} hit rate {tot_hits/tot_pos:.4f}  fix-ups per read {missing/rd.numReads:.3f}  "
        f"phase-2 iterations per tile: max lane {tot_max/tiles:.1f}, equal shares {tot_mean/tiles:.1f} (ratio {tot_max/tot_mean:.2f})")
