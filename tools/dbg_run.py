import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200.utils import synth
from h2gcn_b200 import _cabi
dev=torch.device('cuda:0')
adj=synth.uniform_graph(10000,200000,seed=0)
g=ShardedGraph(adj,0,1,dev)
x=torch.from_numpy(synth.features(10000,128,0)).to(dev); y=torch.empty(10000,256,device=dev)
for _ in range(3): g.round(x,y,[0,128])
torch.cuda.synchronize()
lib=ctypes.CDLL(_cabi.SO_PATH)
buf=np.zeros(8192,dtype=np.int64)
lib.h2_debug_read(buf.ctypes.data_as(ctypes.c_void_p))
m=buf[:1600].reshape(-1,4)
t0=m[0,0]
prev=t0
for i in range(0,180):
    r=m[i]
    if r[0]==0: break
    print(i, "top", int(r[0]-t0), "period", int(r[0]-prev), "wait", int(r[1]-r[0]), "issue", int(r[2]-r[1]), "commit", int(r[3]-r[2]))
    prev=r[0]
