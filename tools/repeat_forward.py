"""Repeat the same forward (same frames, same mask split, same weights) and report how far the six loss terms move:
float atomics give ~1e-7 relative noise; anything larger is a race.  Single process; `--steps` full train steps with
lr=0 (so the weights never change) exercise backward and the optimiser path as well."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomae_b200 as G
from geomae_b200.registry import Config
from geomae_b200.synthetic import make_frame
from geomae_b200.train import FlatTrainer
dev = torch.device("cuda:0")
cfg = Config.fromfile(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs/mae_sst/geomae_nus_pretrain.py"))
torch.manual_seed(0)
model = G.build_detector(cfg.model).to(dev); model.set_impl(sys.argv[1] if len(sys.argv) > 1 else "tc3"); model.train()
model.keep_targets = True
tr = FlatTrainer(model, lr=0.0, weight_decay=0.0)
frames = [torch.from_numpy(make_frame(60 + s, point_scale=0.5)).to(dev) for s in range(3)]
torch.manual_seed(1)
tr.train_step(frames)
ids = (model.last_targets["ids_keep"].clone(), model.last_targets["ids_mask"].clone())
torch.cuda.synchronize()
rows, grads = [], []
N = int(sys.argv[2]) if len(sys.argv) > 2 else 60
names = [k for k, _ in tr.order]
for i in range(N):
    loss, parts = tr.train_step(frames, ids=ids)
    if i < 12:
        grads.append(tr.flat_grad.clone())          # lr = 0: the gradient of the same step, again and again
    rows.append([float(v) for v in model.last_loss_vector.tolist()] if getattr(model, "last_loss_vector", None) is not None else [float(loss)])
rows = np.array(rows)
med = np.median(rows, axis=0)
rel = np.abs(rows - med) / np.abs(med)
print("loss terms (median):", np.round(med, 5))
print("max relative deviation per term:", rel.max(axis=0))
bad = np.where(rel.max(axis=1) > 2e-6)[0]
print("steps deviating by more than 2e-6:", bad.tolist(), [rel[j].max() for j in bad])

# gradients: relative deviation of every parameter's gradient from the first repeat (normal-regression noise reaches the
# density decoder and, through the shared encoder, everything upstream at a much smaller level)
off, worst = 0, []
ref = grads[0]
for k, p in tr.order:
    n = p.numel()
    a = ref[off:off + n]
    dev_ = max(float((g[off:off + n] - a).norm() / (a.norm() + 1e-20)) for g in grads[1:])
    worst.append((dev_, k))
    off += (n + 63) // 64 * 64
worst.sort(reverse=True)
print("largest relative gradient deviation between repeats:", [(f"{d:.2e}", k) for d, k in worst[:5]])
print("median over parameters:", f"{np.median([d for d, _ in worst]):.2e}")
