import faulthandler, os, sys, time
faulthandler.dump_traceback_later(45, exit=True)
import numpy, torch, torch.distributed as dist
rank=int(os.environ['RANK']); local=int(os.environ['LOCAL_RANK']); world=int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
sys.path.insert(0, '.')
from wendy_b200 import multi
from bench import sech2_ic
n=int(float(sys.argv[1])) if len(sys.argv)>1 else 200000
x,v,m=sech2_ic(n, 2+rank)
comm=multi.TorchComm(device='cuda')
ids=(numpy.arange(n)+rank*n).astype(numpy.int32)
m0=1./(n*world)
s=multi.ShardedSystem(x,v,ids,m0,m0*n*world,comm,omega=1.1)
t=time.time()
for i in range(3):
    s.step(1e-3, 5)
    torch.cuda.synchronize()
    print(rank, 'call', i, 'ok %.2fs'%(time.time()-t), 'counts', s.counts.tolist(), 'migrated', s.migrated, flush=True)
s.close()
dist.destroy_process_group()
print(rank,'done',flush=True)
