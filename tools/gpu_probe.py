import sys, os, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import conftest
pkg = conftest.load_package()
P = pkg.semantickitti_params()
orc = conftest.Oracle(P)
ssc = pkg.SSC(P, max_points=131072, max_batch=8)
print('grid', ssc.range_num, ssc.sector_num, ssc.azimuth_num, ssc.bin_num, orc.grid_dims())
scans=[]; poses=[]
for k in range(4):
    s,p = pkg.synth_scan(conftest.SEED, k); scans.append(s); poses.append(p)
poses=np.stack(poses)
# atan2f
rng=np.random.default_rng(1)
y=rng.uniform(-80,80,2_000_000).astype(np.float32); x=rng.uniform(-80,80,2_000_000).astype(np.float32)
d=ssc.atan2f_device(y,x); h=orc.atan2f_many(y,x)
print('atan2f mismatches', int((d.view(np.uint32)!=h.view(np.uint32)).sum()))
# bin
b_g=ssc.makeApriVec(scans[0]); b_o=orc.bin(scans[0])
for k in b_g: print('bin',k,int((b_g[k].view(np.uint32 if b_g[k].dtype==np.float32 else b_g[k].dtype)!=b_o[k].view(np.uint32 if b_o[k].dtype==np.float32 else b_o[k].dtype)).sum()))
# ground
t=time.time(); g,ng=ssc.extractGroudByPatchWork(scans[0]); print('ground gpu %.3f s'%(time.time()-t))
og,ong,ocls,orec=orc.ground(scans[0])
print('ground sizes', len(g),len(og),len(ng),len(ong),'equal', np.array_equal(g,og), np.array_equal(ng,ong))
if not np.array_equal(ng,ong):
    print(' set-equal ng', set(ng.tolist())==set(ong.tolist()), 'set-equal g', set(g.tolist())==set(og.tolist()))
    rec=ssc.last_patch_records(0)
    # compare patch records
    bad=0
    for r in orec:
        z,ri,se=int(r[0]),int(r[1]),int(r[2]); base=[0,32,160,376][z]; secs=[16,32,54,32][z]; pid=base+ri*secs+se
        gr=rec[pid]
        if not (np.array_equal(gr[0:3].view(np.uint32),r[5:8].view(np.uint32)) and int(gr[10])==int(r[4])):
            bad+=1
            if bad<6: print(' patch',pid,'n',r[3],gr[11],'normal',gr[0:3],r[5:8],'mean',gr[3:6],r[8:11],'sv',gr[6:9],r[11:14],'dec',gr[10],r[4])
    print(' bad patches',bad,'of',len(orec))
# full
t=time.time(); ssc.process(scans); print('process gpu %.3f s'%(time.time()-t))
for s in scans: orc.push_scan(s)
for f in range(len(scans)):
    print('counts gpu',ssc.frame_counts(f).tolist(),'orc',orc.counts(f).tolist())
    a_src,a_vid=ssc.frame_apri(f); o_src,o_vid=orc.apri(f)
    print(' apri equal', np.array_equal(a_src,o_src), np.array_equal(a_vid,o_vid))
    vg=ssc.frame_voxels(f); vo=orc.voxels(f)
    for k in vg:
        if vg[k].shape!=vo[k].shape: print('  vox',k,'shape',vg[k].shape,vo[k].shape); continue
        if vg[k].dtype==np.float32: print('  vox',k,'maxabs',float(np.max(np.abs(vg[k]-vo[k]))) if vg[k].size else 0,'bitdiff',int((vg[k].view(np.uint32)!=vo[k].view(np.uint32)).sum()))
        else: print('  vox',k,'diff',int((vg[k]!=vo[k]).sum()))
    for st in range(3):
        print('  stage',st,'names equal', np.array_equal(ssc.frame_point_cluster(f,st), orc.point_cluster(f,st)))
    cg=ssc.frame_clusters(f); co=orc.clusters(f)
    print('  clusters order equal', np.array_equal(cg['name'],co['name']), 'type', np.array_equal(cg['type'],co['type']), 'npts',np.array_equal(cg['npts'],co['npts']),'bbox',np.array_equal(cg['bbox'],co['bbox']))
t=time.time(); ssc.tracking(poses); print('track gpu %.3f s'%(time.time()-t)); orc.track(poses)
for f in range(len(scans)):
    lg=ssc.frame_labels(f); lo=orc.labels(f)
    print('labels frame',f,'mismatch',int((lg!=lo).sum()),'hist',np.bincount(lg,minlength=8).tolist(), np.bincount(lo,minlength=8).tolist())
    cg=ssc.frame_clusters(f); co=orc.clusters(f)
    print('  clusters after track: name',np.array_equal(cg['name'],co['name']),'state',np.array_equal(cg['state'],co['state']),'type',np.array_equal(cg['type'],co['type']))
print('launches', ssc.kernel_launches)
