"""torchrun --nproc-per-node 2 tools/peer_stress.py : the peer all-reduce under rank skew (random device-side delays),
checked against NCCL for every call."""
import datetime, os, sys, random
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
from geomae_b200.peer import PeerExchange
px = PeerExchange.get(dev)
g = torch.Generator(device=dev).manual_seed(5 + rank)
rnd = random.Random(17 + rank)
bad = 0
N = 3000
xs = [torch.randn((1, 128, 256, 512)[i % 4], dtype=torch.float64, device=dev, generator=g) for i in range(N)]
got = []
for i in range(N):                                    # peer calls back to back with random skew, no host sync
    if rnd.random() < 0.3:
        torch.cuda._sleep(rnd.randrange(1000, 400000))
    got.append(px.allreduce_(xs[i].clone(), pre_scale=1.0 + rank, post_scale=0.5))
torch.cuda.synchronize()
for i in range(N):
    ref = xs[i] * (1.0 + rank)
    dist.all_reduce(ref)
    ref *= 0.5
    if not torch.equal(ref, got[i]):
        bad += 1
        if bad < 4:
            print(rank, "MISMATCH at call", i, "n", xs[i].numel(), "max err", (ref - got[i]).abs().max().item(), flush=True)
print(rank, "calls", N, "mismatches", bad, "timeout flag", int(px.timeout.item()), flush=True)
dist.destroy_process_group()
