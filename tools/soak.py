"""One-off soak: a GPU batch against the reference engine + oracle gym layer, every observation and reward, full
state at the end.  usage: soak.py [n] [size] [steps] [full]  (full: 120x100 window at the origin)"""
import sys, numpy as np, torch, time
sys.path.insert(0,'.')
from tests.test_mp_env_gpu import _batch, _finish, _compare_step
from oracle.ref_engine import diff_snapshots
a=sys.argv[1:]
n=int(a[0]) if a else 192
size=int(a[1]) if len(a)>1 else 16
steps=int(a[2]) if len(a)>2 else 400
if len(a)>3 and a[3]=='full': size=(120,100)
xs,ys=(0,0) if isinstance(size,tuple) else (16,8)
seeds=np.arange(n)+9000
env,refs=_batch(n,size,True,10**9,seeds,xs=xs,ys=ys)
W,H=refs[0].MAP_X,refs[0].MAP_Y
_finish(env,refs)
rng=np.random.default_rng(123)
t0=time.time()
for t in range(steps):
    acts=rng.integers(0,20*W*H,n).astype(np.int64)
    _compare_step(env,refs,acts,False,t)
bad=0
for i,r in enumerate(refs):
    d=diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i))
    if d: bad+=1; print(i,d[:3])
print("soak done: %d env-ticks, mismatching envs %d, %.0fs"%(n*steps,bad,time.time()-t0))
