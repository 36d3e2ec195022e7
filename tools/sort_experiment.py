import os, sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, make_tiles
from phosphorus_mk2_b200.rays import RayBatch
sc=scenes.terrain(); acc=Accel(sc)
dev=CudaDevice.make(Options(),0); dev.preprocess(sc,acc); dev.upload_scene(sc)
cam=sc.camera; n=cam.film_width*cam.film_height; tiles=make_tiles(cam.film_width,cam.film_height)
F=("px","py","pz","wx","wy","wz","d","u","v","mesh","face","flags")
def part1by2(x):
    x=x.astype(np.uint32)&0x3ff
    x=(x|(x<<16))&0x30000ff; x=(x|(x<<8))&0x300f00f; x=(x|(x<<4))&0x30c30c3; x=(x|(x<<2))&0x9249249
    return x
for which in ("bounce","shadow"):
    pr=dev.device_rays(n); k=dev.wavefront_rays(tiles,pr,which,0,1,42); h=pr.download().slice(0,k); pr.free()
    def bench(hb,label):
        dr=dev.device_rays(hb.n); ms=[]
        for i in range(8):
            dev.flush_l2(); dr.upload(hb); dev.timer_begin(); dev.trace_device(dr); t=dev.timer_end()
            if i>=3: ms.append(t)
        traced=int(((hb.flags&2)==0).sum())
        print(f"{which:7s} {label:28s} {traced/np.mean(ms)/1e3:8.1f} Mrays/s",flush=True); dr.free()
    bench(h,'pipeline order')
    o=np.stack([h.px,h.py,h.pz],1); lo=o.min(0); hi=o.max(0)
    octant=((h.wx<0).astype(np.uint32))|((h.wy<0).astype(np.uint32)<<1)|((h.wz<0).astype(np.uint32)<<2)
    for bits in (0,3,5,7):
        if bits:
            q=((o-lo)/(hi-lo+1e-9)*((1<<bits)-1)).astype(np.uint32)
            m=part1by2(q[:,0])|(part1by2(q[:,1])<<1)|(part1by2(q[:,2])<<2)
        else: m=np.zeros(h.n,np.uint32)
        for name,key in (("octant|morton%d"%bits,(octant.astype(np.uint64)<<32)|m),("morton%d|octant"%bits,(m.astype(np.uint64)<<3)|octant)):
            if bits==0 and name.startswith('morton'): continue
            idx=np.argsort(key,kind='stable'); hs=RayBatch(h.n)
            for f in F: getattr(hs,f)[:]=getattr(h,f)[idx]
            bench(hs,name)
    rng=np.random.default_rng(0); idx=rng.permutation(h.n); hs=RayBatch(h.n)
    for f in F: getattr(hs,f)[:]=getattr(h,f)[idx]
    bench(hs,'shuffled')
