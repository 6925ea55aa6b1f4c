import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import __graft_entry__ as g
pkg=g.load_package()
from importlib import import_module
synth=import_module('limu_b200.synth')
ctx=pkg.Context(0)
scene=synth.Scene(seed=42)
K=230
traj=synth.loop_trajectory(K+1, radius=30.0, step=1.0)
odo=ctx.KissICP(voxel_size=1.0,cap=10,deskew=True,icp_max_iteration=500)
rows=[]
for i in range(K):
    scan=synth.pad_scan(synth.cast_scan(scene,traj[i],traj[i+1],seed=42*100003+i,device='cuda'),128000,seed=i)
    t=time.perf_counter(); odo.register_frame(scan,want_clouds=False); dt=time.perf_counter()-t
    st=odo.stats
    rows.append((i,st.icp.iterations,st.icp.converged,st.n_keypoints,st.n_down,round(st.sigma,3),round(dt*1e6)))
for r in rows[::10]: print(r)
it=np.array([r[1] for r in rows]); print("iters mean first100",it[:100].mean(),"last100",it[-100:].mean(), "max",it.max(), "map", odo.local_map().size())
