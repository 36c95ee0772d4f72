import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from gymcity_b200.engine import EngineBatch
n=int(sys.argv[1]) if len(sys.argv)>1 else 4096
b=EngineBatch(n)
ss=np.arange(n)
b.generate(ss+1000, ss)
rng=np.random.default_rng(0)
def step():
    tools=rng.integers(0,20,n); tools[tools==5]=7
    b.toolDown(tools, rng.integers(0,16,n)+16, rng.integers(0,16,n)+8)
    b.setFunds(2000000); b.controlSimTick()
for i in range(3): step()
b.sync()
for rep in range(2):
    ev0=torch.cuda.Event(enable_timing=True); ev1=torch.cuda.Event(enable_timing=True)
    tot=0
    for i in range(5):
        tools=rng.integers(0,20,n); tools[tools==5]=7
        b.toolDown(tools, rng.integers(0,16,n)+16, rng.integers(0,16,n)+8)
        b.setFunds(2000000)
        ev0.record(); b.controlSimTick(); ev1.record(); torch.cuda.synchronize()
        tot+=ev0.elapsed_time(ev1)
    print('n',n,'tick ms',tot/5,'env-ticks/s',n/(tot/5/1000))
