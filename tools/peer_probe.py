"""torchrun --nproc-per-node 2 tools/peer_probe.py : peer-memory all-reduce vs NCCL, latency of both."""
import datetime, os, sys, time, traceback
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
try:
    from geomae_b200.peer import PeerExchange
    import warnings; warnings.simplefilter("always")
    px = PeerExchange.get(dev)
    print(rank, "exchange:", px is not None, flush=True)
    x = torch.arange(8, dtype=torch.float64, device=dev) + rank
    px.allreduce_(x); torch.cuda.synchronize()
    print(rank, "first allreduce:", x.tolist(), "timeout flag", int(px.timeout.item()), flush=True)
    for name, fn in (("peer kernel", lambda b: px.allreduce_(b)), ("nccl all_reduce", lambda b: dist.all_reduce(b))):
        b = torch.ones(256, dtype=torch.float64, device=dev)
        for _ in range(20): fn(b)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(200): fn(b)
        e1.record(); host = time.perf_counter() - t0; torch.cuda.synchronize()
        if rank == 0: print(f"{name}: device {1e3 * e0.elapsed_time(e1) / 200:.1f} us, host enqueue {1e6 * host / 200:.1f} us per call", flush=True)
except Exception:
    traceback.print_exc()
finally:
    dist.destroy_process_group()
