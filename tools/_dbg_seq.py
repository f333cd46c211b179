import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import bench, pyoracle as o, liorf_b200
N=14
seq = bench.Sequence(N, 0)
gpu = bench.GpuPipeline(seq, 0); gpu.stage(range(N))
cpu = bench.CpuPipeline(seq)
for i in range(N):
    raw,(t0,it,rot,ptr)=seq.frame(i)
    guess = seq.initial_guess(i, cpu.prev)
    fo = gpu.ctx.processFrame(gpu.pin_raw[i].data_ptr(), len(raw), False, t0, it, seq.imu_cols[i], ptr, True, guess, loop_every=10, frame_index=i)
    op = cpu.step(i)
    gp = np.array(fo.pose[:],np.float32)
L = cpu.last
print('frame 13 gpu', gp, 'cpu', op)
g_map = gpu.ctx.getLocalMap(); g_ds = gpu.ctx.getScanDS()
print('map equal', np.array_equal(g_map, L['mds']), 'ds equal', np.array_equal(g_ds, L['ds']))
for name, kw in (('oracle brute', dict(use_ref_kdtree=False)), ('oracle nanoflann', dict(use_ref_kdtree=True))):
    r = o.scan2map(L['ds'], L['mds'], L['guess'], 30, False, L['state'], **kw)
    print(name, 'iters', r['iters'], 'nsel', r['nsel']); print(r['trace'])
c = liorf_b200.Context()
c.setLocalMap(L['mds']); c.setCurrentScan(L['ds']); ds2, n2 = c.downsampleCurrentScan(len(L['ds']))
print('re-downsample identical', np.array_equal(ds2, L['ds']))
c.setLMState(int(L['state'][0]), L['state'][1:])
pose, tr = c.scan2MapOptimization(L['guess'], 30, force_all_iters=False)
print('gpu stepwise iters', tr.iters, 'nsel', tr.nsels()); print(tr.poses())
# per-point comparison at the initial guess
r0 = o.scan2map(L['ds'], L['mds'], L['guess'], 1, True, L['state'])
so = c.surfOptimization(L['guess'], len(L['ds']))
print('gpu flags', int(np.sum(so['flag'])), 'oracle nsel it0', r0['nsel'])

p1 = tr.poses()[0]
so1 = c.surfOptimization(p1, len(L['ds']))
oo = o.surf_optimization(L['ds'], L['mds'], p1); oc, of, oidx, od2, opl = oo['coeff'], oo['flag'], oo['idx'], oo['d2'], oo['plane']
print('hook flags at pose1', int(so1['flag'].sum()), 'oracle', int(np.sum(of)))
diff = np.nonzero(so1['flag'] != of)[0]
print('flag diffs', diff)
print('idx equal', np.array_equal(so1['idx'], oidx), 'coeff equal', np.array_equal(so1['coeff'], oc))
# LM step from the hook's output at pose1 vs the persistent kernel's iteration-1 result
for d_ in diff[:5]:
    print(d_, 'gpu', so1['coeff'][d_], so1['d2'][d_], so1['plane'][d_], 'orc', oc[d_], od2[d_], opl[d_])
