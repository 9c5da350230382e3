import os, sys, time, tomllib, json
ROOT='/root/repo'
sys.path[:0]=[ROOT, ROOT+'/carla-social-force-model_b200']
import numpy as np, torch
from sfm_b200 import engine as eng, synth
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE']); local=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
cfg=tomllib.load(open(ROOT+'/carla-social-force-model_b200/config/sfm_config.toml','rb'))
w=synth.make_config(3, n=92672)
e=eng.Engine(cfg, w.step_length, device=local); e.load(w)
ctx=e.ctx
rows=e.hi-e.lo
pin=lambda: torch.empty((rows,3),dtype=torch.float64).pin_memory()
hl,hv,hnl,hnv=pin(),pin(),pin(),pin()
l0,v0=e.local_state(); hl.numpy()[:]=l0; hv.numpy()[:]=v0
a=[x.numpy() for x in (hl,hv,hnl,hnv)]
for _ in range(5): e.step(1,True)
def barrier():
    torch.cuda.synchronize(); torch.distributed.barrier(); torch.cuda.synchronize()
res={}
for label in ('tick_host','pieces'):
    ctx.reset_stats(); ctx.set_profiling(True)
    ts=[]
    for k in range(8):
        barrier()
        t0=time.perf_counter()
        if label=='tick_host':
            e.tick_host(a[0],a[1],a[3],a[2])
            torch.cuda.synchronize()
            ts.append((time.perf_counter()-t0)*1e3)
        else:
            ctx.update_kinematics(a[0],a[1]); t1=time.perf_counter()
            ctx.stage(); torch.cuda.synchronize(); t2=time.perf_counter()
            ctx.step_peer(1,True); torch.cuda.synchronize(); t3=time.perf_counter()
            ctx.download_state(a[2],a[3]); t4=time.perf_counter()
            ts.append([(t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3])
        a[0][...]=a[2]; a[1][...]=a[3]
    s=ctx.stats(); ctx.set_profiling(False)
    res[label]=dict(ts=np.round(np.array(ts[2:]).mean(0),3).tolist(), ms_pairs=s['ms_pairs']/8, ms_seg=s['ms_segments']/8, ms_cells=s['ms_cells']/8, ms_int=s['ms_integrate']/8)
if rank==0: print(os.environ.get('SFM_K2_PERSIST_MULTI','0'), json.dumps(res))
torch.distributed.barrier(); torch.distributed.destroy_process_group()
