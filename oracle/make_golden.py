"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the UNMODIFIED reference.

Run here (the build container, where /root/reference exists):

    python -m oracle.make_golden

For each case the reference detector is built from its own config
(configs/mae_sst/…6x_1e-5.py), its parameters are overwritten with
``geomae_oracle.init_params(cfg, seed)`` (so a fixture only has to carry the
seed, not 2.7 M weights), the keep/mask split is injected from a seeded CPU
generator, and ``forward_train`` + ``backward`` run on CPU.  Captured tensors
are written to ``tests/golden/<case>.npz``.
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from geomae_b200.synthetic import make_frame  # noqa: E402
from oracle import geomae_oracle as O  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# geometries other than the reference config's (BASELINE.json configs[3] / configs[4]): the reference code is generic
# in range / voxel sizes / grid (it reads them from the config), so the UNMODIFIED detector is built from its own
# config with only those values replaced (apply_geometry)
WAYMO_GEOMETRY = dict(pc_range=(-74.88, -74.88, -2.0, 74.88, 74.88, 4.0), voxel_size=(0.32, 0.32, 6),
                      sub_voxel_size_med=(0.16, 0.16, 1.5), sub_voxel_size_low=(0.08, 0.08, 0.75),
                      grid_size=(1, 468, 468))
DENSE_GEOMETRY = dict(pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), voxel_size=(0.1, 0.1, 8),
                      sub_voxel_size_med=(0.05, 0.05, 2), sub_voxel_size_low=(0.025, 0.025, 1),
                      grid_size=(1, 1024, 1024))

CASES = {
    # name: (frame kwargs per sample, enc blocks, dec blocks, store full tensors?)
    "small_b2": dict(frames=[dict(seed=11, point_scale=0.08), dict(seed=12, point_scale=0.04, sweeps=6)],
                     enc=1, dec=1, full=True, param_seed=3, mask_seed=5),
    "config0_1frame_1block": dict(frames=[dict(seed=21)], enc=1, dec=1, full=False,
                                  param_seed=4, mask_seed=6),
    "full_b2": dict(frames=[dict(seed=31), dict(seed=32)], enc=6, dec=2, full=False,
                    param_seed=7, mask_seed=8),
    # BASELINE.json configs[3]: Waymo-shaped frames (64 beams), 0.32 m pillars, 468 x 468 grid
    "waymo_b2": dict(frames=[dict(seed=41, preset="waymo", point_scale=0.12), dict(seed=42, preset="waymo", point_scale=0.06)],
                     enc=1, dec=1, full=False, param_seed=9, mask_seed=10, geometry=WAYMO_GEOMETRY),
    # BASELINE.json configs[4]: dense-grid stress, 0.1 m pillars on the nuScenes range (1024 x 1024 grid), multi-sweep
    "dense_b1": dict(frames=[dict(seed=51, point_scale=0.25, sweeps=3)], enc=1, dec=1, full=False,
                     param_seed=11, mask_seed=12, geometry=DENSE_GEOMETRY),
}


def case_frames(case):
    return [make_frame(**kw) for kw in case["frames"]]


def case_cfg(case):
    return O.PathConfig(enc_blocks=case["enc"], dec_blocks=case["dec"], **case.get("geometry", {}))


def apply_geometry(model_cfg, geometry, base=None):
    """Copy of a model config dict (the reference's or ours: same keys) with range, the three voxel sizes, grid_size,
    spatial_shape and output_shape replaced.  Voxel sizes are recognised by VALUE (the config repeats them in seven sub-dicts)."""
    base = base or O.PathConfig()
    sizes = {tuple(base.voxel_size): tuple(geometry["voxel_size"]),
             tuple(base.sub_voxel_size_med): tuple(geometry["sub_voxel_size_med"]),
             tuple(base.sub_voxel_size_low): tuple(geometry["sub_voxel_size_low"])}
    gz, gy, gx = geometry["grid_size"]

    def walk(node):
        if isinstance(node, dict):
            out = type(node)()
            for k, v in node.items():
                if k == "point_cloud_range":
                    out[k] = list(geometry["pc_range"])
                elif k in ("voxel_size", "sub_voxel_size_low", "sub_voxel_size_med", "sub_voxel_size_top") and \
                        tuple(v) in sizes:
                    out[k] = sizes[tuple(v)]
                elif k == "grid_size":
                    out[k] = (gz, gy, gx)
                elif k == "spatial_shape":
                    out[k] = [gz, gy, gx]
                elif k == "output_shape":
                    out[k] = [gy, gx]
                else:
                    out[k] = walk(v)
            return out
        if isinstance(node, (list, tuple)):
            return type(node)(walk(v) for v in node)
        return node
    return walk(model_cfg)


def run_reference(case):
    frames = case_frames(case)
    cfg = case_cfg(case)
    model_cfg = None
    if "geometry" in case:
        model_cfg = apply_geometry(H.load_config_model(), case["geometry"])
    det = H.build_detector(model_cfg, seed=0, encoder_blocks=case["enc"], decoder_blocks=case["dec"])
    params = O.init_params(cfg, case["param_seed"])
    sd = det.state_dict()
    missing = [k for k in params if k not in sd]
    assert not missing, missing
    det.load_state_dict({**sd, **params})
    # the split needs the sorted pillar list; take it from the reference's own VFE output order
    coors_top = O.batch_voxelize(frames, cfg.voxel_size, cfg.pc_range)
    rows, _, _ = O.unique_rows(coors_top)
    keep, mask = O.vanilla_mask_ids(rows, len(frames), cfg.mask_ratio, case["mask_seed"])
    rec = H.Recorder(det, ids=(torch.from_numpy(keep), torch.from_numpy(mask)))
    pts = [torch.from_numpy(f) for f in frames]
    losses = det.forward_train(points=pts, img_metas=[{} for _ in pts])
    total = sum(losses.values())
    total.backward()
    rec.close()
    grads = {k: p.grad for k, p in det.named_parameters()}
    return frames, keep, mask, rec.out, losses, grads, det


def pack(case_name, case):
    frames, keep, mask, out, losses, grads, det = run_reference(case)
    g = {}
    g["points_crc"] = np.array([zlib.crc32(f.tobytes()) for f in frames], np.int64)
    g["n_points"] = np.array([f.shape[0] for f in frames], np.int64)
    g["ids_keep"], g["ids_mask"] = keep, mask
    for k, v in losses.items():
        g["loss/" + k] = np.float64(v.item())
    for k, v in grads.items():
        g["gradnorm/" + k] = np.float64(v.double().norm().item())
    feats, fcoors = out["voxel_encoder"]
    g["n_pillars"] = np.int64(fcoors.shape[0])
    g["voxel_features_absum"] = np.float64(feats.double().abs().sum().item())
    if case["full"]:
        for i, f in enumerate(frames):
            g[f"points{i}"] = f
        g["coors_top"] = out["voxelize"][1].numpy()
        g["coors_low"] = out["sub_voxelize_low"].numpy()
        g["coors_med"] = out["sub_voxelize_med"].numpy()
        g["pillar_coors"] = fcoors.numpy()
        g["voxel_features_rows8"] = feats.detach().numpy()[::8]
        for name, (cen, rows, cnt) in zip(("low", "med", "top"), out["get_centroid_per_voxel"]):
            g[f"centroid_{name}"] = cen.numpy()
            g[f"rows_{name}"] = rows.numpy()
            g[f"count_{name}"] = cnt.numpy()
        med_raw, med_mask = out["get_multi_voxel_id_to_tensor_id_for_curv"]
        g["med_raw_vals"], g["med_mask"] = med_raw.numpy()[med_mask.numpy()], med_mask.numpy()
        g["pair"] = out["pair"].numpy()
        normal, curv = out["cal_regular_voxel_nor_and_curv"]
        g["normal"], g["curvature"] = normal.numpy(), curv.numpy()
        tl, tlm, tm, tmm = out["get_multi_voxel_id_to_tensor_id_ori"]
        g["tgt_low_mask"], g["tgt_med_mask"] = tlm.numpy(), tmm.numpy()
        g["tgt_low_vals"] = tl.numpy()[tlm.numpy()]
        g["tgt_med_vals"] = tm.numpy()[tmm.numpy()]
        x = out["extract_feat"]
        g["tgt_top"], g["tgt_normal"] = x[5].numpy(), x[6].numpy()
        pred = out["backbone"]
        g["pred_reg_top"] = pred[2].detach().numpy()
        g["pred_nor_top"] = pred[5].detach().numpy()
        g["pred_reg_med"] = pred[1].detach().numpy()
        g["pred_cls_med"] = pred[7].detach().numpy()
        for k in ("voxel_encoder.vfe_layers.0.linear.weight", "voxel_encoder.vfe_layers.1.norm.weight",
                  "backbone.mask_token", "backbone.decoder_pred_top.weight",
                  "backbone.encoder_blocks.0.encoder_list.0.win_attn.self_attn.in_proj_bias",
                  "backbone.encoder_blocks.0.encoder_list.1.norm2.weight"):
            g["grad/" + k] = grads[k].numpy()
        bb = det.backbone
        # window layout of the decoder token set, shift 1 (reference's own bookkeeping)
        vis = fcoors[torch.from_numpy(keep)]
        msk = fcoors[torch.from_numpy(mask)]
        info = bb.window_partition(torch.cat([vis, msk]).long(), {})
        info = bb.get_voxel_keep_inds(info, 2)
        for s in (0, 1):
            g[f"dec_win_shift{s}"] = info[f"batch_win_inds_shift{s}"].numpy()
            g[f"dec_lvl_shift{s}"] = info[f"voxel_drop_level_shift{s}"].numpy()
            g[f"dec_ciw_shift{s}"] = info[f"coors_in_win_shift{s}"].numpy()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, case_name + ".npz")
    np.savez_compressed(path, **g)
    print(case_name, "->", path, f"{os.path.getsize(path)/1024:.0f} KiB",
          {k: round(float(v), 4) for k, v in losses.items()})


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        pack(n, CASES[n])
