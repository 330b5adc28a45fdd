import os, sys, torch, cProfile, pstats
sys.path.insert(0, "/root/repo")
import geomae_b200
from geomae_b200.registry import Config, build_model
from geomae_b200.synthetic import make_frame
from geomae_b200.train import FlatTrainer
dev = torch.device("cuda:0")
cfg = Config.fromfile("/root/repo/configs/mae_sst/geomae_nus_pretrain.py")
model = build_model(cfg.model).to(dev).train(); model.set_impl("tc1")
tr = FlatTrainer(model)
frames = [torch.from_numpy(make_frame(s + 1, point_scale=0.02)).to(dev) for s in range(4)]
for _ in range(3): tr.train_step(frames)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): tr.train_step(frames)
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats("geomae_b200|built-in method torch|method .* of .torch", 60)
print("=" * 30, "by self time")
st.sort_stats("tottime").print_stats(45)
