import os, sys, time, torch, torch.distributed as dist
rank=int(os.environ['RANK']); local=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
t=time.time()
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
x=torch.ones(4, device='cuda')*rank
dist.all_reduce(x); torch.cuda.synchronize()
print(rank, 'allreduce ok', x.tolist(), 'in %.1fs'%(time.time()-t), flush=True)
# p2p batch
ops=[]
peer=1-rank
r=torch.empty(3, device='cuda'); s=torch.full((3,), float(rank), device='cuda')
ops=[dist.P2POp(dist.irecv, r, peer), dist.P2POp(dist.isend, s, peer)]
for q in dist.batch_isend_irecv(ops): q.wait()
torch.cuda.synchronize()
print(rank, 'p2p ok', r.tolist(), 'in %.1fs'%(time.time()-t), flush=True)
sys.path.insert(0, '.')
from wendy_b200 import multi
c=multi.TorchComm(device='cuda')
print(rank, 'allgather_vec', c.allgather_vec([rank, 5]).tolist(), flush=True)
send=[torch.zeros((0,3),dtype=torch.float64,device='cuda') for _ in range(2)]
send[peer]=torch.full((rank+2,3), float(rank), dtype=torch.float64, device='cuda')
rec=c.exchange(send)
print(rank, 'exchange ok', [tuple(t.shape) for t in rec], 'in %.1fs'%(time.time()-t), flush=True)
dist.destroy_process_group()
