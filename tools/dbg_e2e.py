import os, sys, time, torch
sys.path.insert(0, '/root/repo')
import geomae_b200
from geomae_b200.registry import Config, build_model
from geomae_b200.synthetic import make_frame
from geomae_b200.train import FlatTrainer
dev = torch.device("cuda:0")
cfg = Config.fromfile('/root/repo/configs/mae_sst/geomae_nus_pretrain.py')
model = build_model(cfg.model).to(dev).train(); model.set_impl("tc1")
tr = FlatTrainer(model)
host = [[torch.from_numpy(make_frame(16*b + s + 1)).pin_memory() for s in range(4)] for b in range(4)]
res = [[f.to(dev) for f in b] for b in host]
def loop(fn, n=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return 1e3*(time.perf_counter()-t0)/n
print("resident            ", loop(lambda i: tr.train_step(res[i % 4])))
print("resident + float    ", loop(lambda i: float(tr.train_step(res[i % 4])[0])))
print("host                ", loop(lambda i: tr.train_step_from_host(host[i % 4])))
print("host + float        ", loop(lambda i: float(tr.train_step_from_host(host[i % 4])[0])))
print("resident again      ", loop(lambda i: tr.train_step(res[i % 4])))
