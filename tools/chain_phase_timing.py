#!/usr/bin/env python
"""Phase stamps (clock64 of CTA 0, first tile) of k_sra_chain_fwd: where a 128-token tile spends its cycles.
usage (GPU box): python tools/chain_phase_timing.py [n_tokens]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from geomae_b200 import lib as L  # noqa: E402
from geomae_b200.dense import sra_chain_fwd  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 58 * 128
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)  # noqa: E731
lay = dict(Wo=rnd(128, 128, sc=0.1), bo=rnd(128, sc=0.1), W1=rnd(256, 128, sc=0.1), b1=rnd(256, sc=0.1),
           W2=rnd(128, 256, sc=0.1), b2=rnd(128, sc=0.1), g1=1 + 0.1 * rnd(128), be1=0.1 * rnd(128),
           g2=1 + 0.1 * rnd(128), be2=0.1 * rnd(128), eps=1e-5)
nxt = (rnd(384, 128, sc=0.1), rnd(384, sc=0.1))
x, attn = rnd(n, 128), rnd(n, 128).to(torch.bfloat16)
table, cell = rnd(144, 128), torch.randint(0, 144, (n,), dtype=torch.int32, device=dev)
stamps = torch.zeros(32, dtype=torch.int64, device=dev)
fn = L.lib().geomae_debug_chain_stamps
fn.argtypes, fn.restype = [C.c_void_p], C.c_int
names = ["body start", "inputs on chip", "out-proj acc", "E1 (LN1) done", "FFN1 acc", "E2 (GELU) done", "FFN2 acc",
         "E3 (LN2) done", "next-layer operands", "in-proj acc", "tile done (E4)"]
for rep in range(3):
    fn(stamps.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sra_chain_fwd(x, attn=attn, layer=lay, next_in_proj=nxt, pos_table=table, tok_cell_next=cell)
    e1.record()
    torch.cuda.synchronize()
    fn(None)
    t = stamps.cpu().tolist()
    print(f"rep {rep}: n={n} kernel+pack {e0.elapsed_time(e1) * 1e3:.1f} us (includes 4 weight-pack launches)")
    for i in range(1, len(names)):
        print(f"  {names[i]:18s} +{t[i] - t[i - 1]:7d} cycles   (t = {t[i] - t[0]:7d})")
