"""torchrun --nproc-per-node N tools/exp/upload_prof.py : times the phases of RankSlab.upload / download_owned."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import numpy as np, torch, torch.distributed as dist
from openabl_b200.model import Model
from openabl_b200.slab import RankSlab, exchange_partitions
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
m = Model(os.path.join(REPO, "examples", "boids2d.abl"), {"num_agents": 1000000 * world})
m.populate()
slab = RankSlab(m, rank, world, dist, device=lr)
slab.upload(); slab.upload()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    dist.barrier(); t0 = T()
    arr = m.host_view(0); n = len(arr); pool = m.pool(0)
    lo, hi = n * rank // world, n * (rank + 1) // world
    rec = slab.rt.transit_record_bytes(pool)
    part = arr[lo:hi]
    send = torch.empty((hi - lo) * rec, dtype=torch.uint8, device="cuda")
    t1 = T()
    counts = slab.rt.partition_upload(pool, part, lo, send.data_ptr(), world)
    t2 = T()
    recv, total = exchange_partitions(dist, rank, world, send, counts, rec)
    t3 = T()
    slab.rt.adopt_records(pool, recv.data_ptr(), total, n)
    t4 = T()
    slab.rt.exchange(pool); slab.rt.synchronize()
    t5 = T()
    for _ in range(20): slab.timestep()
    slab.rt.synchronize(); t6 = T()
    out = slab.download_owned(0)
    t7 = T()
    if rank == 0:
        print("rep %d: alloc %.2f partition_upload %.2f all2all %.2f adopt %.2f exchange %.2f 20 steps %.2f download %.2f | total %.2f ms" % (
            rep, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), 1e3*(t6-t5), 1e3*(t7-t6), 1e3*(t7-t0)), flush=True)
m.close()
dist.destroy_process_group()
