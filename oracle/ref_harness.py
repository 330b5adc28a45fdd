"""TEST INFRASTRUCTURE — runs the *unmodified* GeoMAE reference on CPU.

Only usable where ``/root/reference`` exists (the build container).  It is the
generator behind ``tests/golden/*.npz`` (see ``oracle/make_golden.py``) and the
thing ``oracle/geomae_oracle.py`` (the portable restatement) is pinned against.
Nothing in the product package imports this file.

How it works: the reference cannot be imported as a
package (mmcv/mmdet/mmseg/spconv/torch_scatter/ipdb are not installed), so we
  1. compile the reference's own CPU voxelisation sources
     (``mmdet3d/ops/voxel/src/{voxelization.cpp,voxelization_cpu.cpp,
     scatter_points_cpu.cpp}``) in place into ``oracle/_ref/``,
  2. register small stub modules for the absent third-party packages,
  3. load the reference's hot-path files *by path* under their real dotted
     names, and build the detector from the reference's own config dict.

The only arithmetic here that is not executed from reference source are the
three third-party pieces that are absent from ``/root/reference``:
``torch_scatter.scatter/scatter_max`` (unpinned), spconv 2.1.21
``get_indice_pairs_implicit_gemm`` (3x3 sub-manifold neighbour table) and mmdet
2.20.0 ``CrossEntropyLoss(use_sigmoid=True)`` / ``SmoothL1Loss``; they are
restated below from their published behaviour.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("GEOMAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(HERE, "_ref")

MAE_CONFIG = ("configs/mae_sst/"
              "m_sst_nus_singlestage_curv_07_ssl_dataset_wo_dbsampler_6x_1e-5.py")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mmdet3d"))


# --------------------------------------------------------------------------
# 1. the reference's own C++ voxelisation, compiled where it lies
# --------------------------------------------------------------------------
def build_voxel_layer():
    """Compile ``voxel_layer`` (CPU build: WITH_CUDA undefined) into oracle/_ref."""
    from torch.utils.cpp_extension import load
    os.makedirs(REF_BUILD, exist_ok=True)
    src = os.path.join(REF_ROOT, "mmdet3d/ops/voxel/src")
    return load(
        name="voxel_layer_ref",
        sources=[os.path.join(src, f) for f in
                 ("voxelization.cpp", "voxelization_cpu.cpp", "scatter_points_cpu.cpp")],
        extra_cflags=["-O2", "-w"],
        build_directory=REF_BUILD,
        verbose=False,
    )


def load_prebuilt_voxel_layer():
    """Import oracle/_ref/voxel_layer_ref.so without the sources (GPU box)."""
    so = os.path.join(REF_BUILD, "voxel_layer_ref.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("voxel_layer_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# --------------------------------------------------------------------------
# 2. stubs for absent third-party packages
# --------------------------------------------------------------------------
class Registry:
    """mmcv.utils.Registry, reduced to register_module()/build()."""

    def __init__(self, name, parent=None, **_):
        self.name = name
        self._mods = parent._mods if parent is not None else {}

    def register_module(self, name=None, force=False, module=None):
        if isinstance(name, type):  # bare decorator
            self._mods[name.__name__] = name
            return name

        def deco(cls):
            self._mods[name or cls.__name__] = cls
            return cls
        return deco

    def get(self, key):
        return self._mods.get(key)

    def build(self, cfg, default_args=None):
        cfg = dict(cfg)
        for k, v in (default_args or {}).items():
            cfg.setdefault(k, v)
        return self._mods[cfg.pop("type")](**cfg)


def _identity_decorator(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return lambda fn: fn


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, **attrs):
    m = _mod(name, **attrs)
    m.__path__ = []
    return m


def scatter_mean_restated(src, index, dim=0, reduce="mean"):
    """torch_scatter.scatter(reduce='mean'|'sum'): sum per index, divide by count."""
    assert dim == 0
    n = int(index.max()) + 1
    out = src.new_zeros((n,) + src.shape[1:])
    out.index_add_(0, index, src)
    if reduce == "sum":
        return out
    cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))


class _ScatterMax(torch.autograd.Function):
    """torch_scatter.scatter_max: per-index max; grad goes to one arg-max row
    (smallest point index among ties, as the in-repo op does,
    reference mmdet3d/ops/voxel/src/scatter_points_cuda.cu:154-158)."""

    @staticmethod
    def forward(ctx, src, index):
        n = int(index.max()) + 1
        out = torch.full((n, src.shape[1]), float("-inf"), dtype=src.dtype)
        out = out.scatter_reduce(0, index.view(-1, 1).expand_as(src), src, "amax", include_self=True)
        is_max = src == out[index]
        rows = torch.arange(src.shape[0]).view(-1, 1).expand_as(src)
        big = src.shape[0]
        cand = torch.where(is_max, rows, torch.full_like(rows, big))
        arg = torch.full((n, src.shape[1]), big, dtype=torch.long)
        arg = arg.scatter_reduce(0, index.view(-1, 1).expand_as(src), cand, "amin", include_self=True)
        ctx.save_for_backward(arg)
        ctx.n_src = src.shape[0]
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g_out, _g_arg):
        (arg,) = ctx.saved_tensors
        g = g_out.new_zeros((ctx.n_src, g_out.shape[1]))
        g.scatter_(0, arg, g_out)
        return g, None


def scatter_max_restated(src, index, dim=0):
    assert dim == 0
    return _ScatterMax.apply(src, index)


def subm_pairs_3x3(indices, batch_size, spatial_shape):
    """spconv 2.x get_indice_pairs_implicit_gemm(subm=True, ksize=[1,3,3]) -> pair [9, N].

    pair[k, i] = row of the active voxel at offset k=(dy+1)*3+(dx+1) from voxel i,
    -1 when that cell is empty / outside the grid."""
    idx = indices.long()
    _, ny, nx = spatial_shape
    n = idx.shape[0]
    table = torch.full((batch_size * ny * nx,), -1, dtype=torch.long)
    table[idx[:, 0] * ny * nx + idx[:, 2] * nx + idx[:, 3]] = torch.arange(n)
    pair = torch.full((9, n), -1, dtype=torch.int32)
    k = 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            y, x = idx[:, 2] + dy, idx[:, 3] + dx
            ok = (y >= 0) & (y < ny) & (x >= 0) & (x < nx)
            key = idx[:, 0] * ny * nx + y.clamp(0, ny - 1) * nx + x.clamp(0, nx - 1)
            pair[k] = torch.where(ok, table[key], torch.full_like(key, -1)).int()
            k += 1
    return pair


class SmoothL1LossRestated(nn.Module):
    def __init__(self, beta=1.0, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.beta, self.reduction, self.loss_weight = beta, reduction, loss_weight

    def forward(self, pred, target, **_):
        return self.loss_weight * F.smooth_l1_loss(pred, target, beta=self.beta, reduction=self.reduction)


class CrossEntropyLossRestated(nn.Module):
    """mmdet 2.20 CrossEntropyLoss(use_sigmoid=True): one-hot expand the labels to
    pred's channel count, BCE-with-logits, mean over every element."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction="mean",
                 class_weight=None, ignore_index=None, loss_weight=1.0):
        super().__init__()
        assert use_sigmoid and reduction == "mean"
        self.loss_weight = loss_weight

    def forward(self, cls_score, label, **_):
        onehot = F.one_hot(label, cls_score.shape[-1]).to(cls_score.dtype)
        return self.loss_weight * F.binary_cross_entropy_with_logits(cls_score, onehot, reduction="mean")


_INSTALLED = {}


def install(voxel_layer=None):
    """Register stubs + load the reference hot-path files. Idempotent."""
    if _INSTALLED:
        return _INSTALLED
    assert available(), f"reference tree not found at {REF_ROOT}"
    voxel_layer = voxel_layer or build_voxel_layer()

    regs = {n: Registry(n) for n in
            ("DETECTORS", "BACKBONES", "HEADS", "NECKS", "LOSSES", "ROI_EXTRACTORS",
             "SHARED_HEADS", "SEGMENTORS", "NORM_LAYERS", "MODELS")}
    regs["LOSSES"].register_module("SmoothL1Loss")(SmoothL1LossRestated)
    regs["LOSSES"].register_module("CrossEntropyLoss")(CrossEntropyLossRestated)
    for n in ("BN", "BN1d"):
        regs["NORM_LAYERS"].register_module(n)(nn.BatchNorm1d)

    def build_norm_layer(cfg, num_features, postfix=""):
        cfg = dict(cfg)
        cls = regs["NORM_LAYERS"].get(cfg.pop("type"))
        cfg.pop("requires_grad", None)
        return "bn" + str(postfix), cls(num_features, **cfg)

    class BaseDetector(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()

        @property
        def with_neck(self):
            return getattr(self, "neck", None) is not None

    _pkg("mmcv")
    _mod("mmcv.runner", force_fp32=_identity_decorator, auto_fp16=_identity_decorator)
    _mod("mmcv.cnn", build_norm_layer=build_norm_layer, build_conv_layer=None,
         NORM_LAYERS=regs["NORM_LAYERS"], MODELS=regs["MODELS"])
    _mod("mmcv.utils", Registry=Registry)
    _mod("mmcv.parallel", DataContainer=object)
    mm_regs = {k: regs[k] for k in ("DETECTORS", "BACKBONES", "HEADS", "NECKS", "LOSSES",
                                   "ROI_EXTRACTORS", "SHARED_HEADS")}
    builders = dict(build_backbone=lambda c: regs["BACKBONES"].build(c),
                    build_head=lambda c: regs["HEADS"].build(c),
                    build_neck=lambda c: regs["NECKS"].build(c))
    _pkg("mmdet")
    _pkg("mmdet.models", **mm_regs, **builders)
    _mod("mmdet.models.builder", **mm_regs, **builders)
    _mod("mmdet.models.detectors", BaseDetector=BaseDetector)
    _pkg("mmseg"); _pkg("mmseg.models")
    _mod("mmseg.models.builder", SEGMENTORS=regs["SEGMENTORS"])
    _mod("ipdb", set_trace=lambda *a, **k: None)
    _mod("torch_scatter", scatter=scatter_mean_restated, scatter_max=scatter_max_restated)
    _pkg("spconv"); _pkg("spconv.pytorch")

    def get_indice_pairs_implicit_gemm(indices, batch_size, spatial_shape, **kw):
        assert kw.get("subm") and list(kw.get("ksize")) == [1, 3, 3]
        pair = subm_pairs_3x3(indices, batch_size, spatial_shape)
        return (None, None, pair, None, None, None, None, None, None)

    _mod("spconv.pytorch.ops", get_indice_pairs=None,
         get_indice_pairs_implicit_gemm=get_indice_pairs_implicit_gemm)
    _mod("spconv.core", ConvAlgo=types.SimpleNamespace(MaskImplicitGemm=0))

    # fake mmdet3d package tree; real files are loaded into it by path below
    for p in ("mmdet3d", "mmdet3d.ops", "mmdet3d.ops.voxel", "mmdet3d.ops.sst", "mmdet3d.models",
              "mmdet3d.models.sst", "mmdet3d.models.voxel_encoders", "mmdet3d.models.backbones",
              "mmdet3d.models.detectors"):
        _pkg(p)
    _mod("mmdet3d.core", bbox3d2result=None, merge_aug_bboxes_3d=None, Box3DMode=None,
         Coord3DMode=None, show_result=None)
    _mod("mmdet3d.ops.voxel.voxel_layer", **{k: getattr(voxel_layer, k) for k in dir(voxel_layer)
                                             if not k.startswith("_")})
    _mod("mmdet3d.models.detectors.voxelnet", VoxelNet=object)
    ops = sys.modules["mmdet3d.ops"]
    ops.spconv = None
    ops.points_in_boxes_cpu = ops.points_in_boxes_gpu = None
    ops.make_sparse_convmodule = None

    def load(dotted, rel):
        spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF_ROOT, "mmdet3d", rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[dotted] = m
        parent, _, leaf = dotted.rpartition(".")
        setattr(sys.modules[parent], leaf, m)
        spec.loader.exec_module(m)
        return m

    vox = load("mmdet3d.ops.voxel.voxelize", "ops/voxel/voxelize.py")
    sp = load("mmdet3d.ops.voxel.scatter_points", "ops/voxel/scatter_points.py")
    ops.Voxelization, ops.Voxelization_with_flag = vox.Voxelization, vox.Voxelization_with_flag
    ops.DynamicScatter = sp.DynamicScatter
    sst = load("mmdet3d.ops.sst.sst_ops", "ops/sst/sst_ops.py")
    for n in ("scatter_v2", "flat2window", "window2flat", "get_inner_win_inds",
              "make_continuous_inds", "get_flat2win_inds"):
        setattr(ops, n, getattr(sst, n))
    load("mmdet3d.ops.norm", "ops/norm.py")
    load("mmdet3d.models.builder", "models/builder.py")
    load("mmdet3d.models.sst.sst_basic_block", "models/sst/sst_basic_block.py")
    load("mmdet3d.models.voxel_encoders.utils", "models/voxel_encoders/utils.py")
    load("mmdet3d.models.voxel_encoders.voxel_encoder", "models/voxel_encoders/voxel_encoder.py")
    load("mmdet3d.models.backbones.multi_mae_sst_spearate_top_only",
         "models/backbones/multi_mae_sst_spearate_top_only.py")
    load("mmdet3d.models.detectors.base", "models/detectors/base.py")
    load("mmdet3d.models.detectors.single_stage", "models/detectors/single_stage.py")
    det = load("mmdet3d.models.detectors.multi_sub_voxel_dynamic_voxelnet_ssl",
               "models/detectors/multi_sub_voxel_dynamic_voxelnet_ssl.py")
    _INSTALLED.update(regs=regs, voxel_layer=voxel_layer, detector_module=det, sst_ops=sst, ops=ops)
    return _INSTALLED


def load_pipeline_loading():
    """The reference's datasets/pipelines/loading.py (LoadPointsFromFile, LoadPointsFromMultiSweeps) loaded by path on
    top of install(): real point classes (core/points/{base,lidar}_points.py), stubbed mmcv file client (disk reads)
    and mmdet pipeline registry."""
    env = install()
    if "loading" in env:
        return env["loading"]

    def load(dotted, rel):
        spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF_ROOT, "mmdet3d", rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[dotted] = m
        spec.loader.exec_module(m)
        return m

    pts_pkg = _pkg("mmdet3d.core.points")
    base = load("mmdet3d.core.points.base_points", "core/points/base_points.py")
    lidar = load("mmdet3d.core.points.lidar_points", "core/points/lidar_points.py")

    def get_points_type(points_type):
        assert points_type == "LIDAR"
        return lidar.LiDARPoints
    pts_pkg.BasePoints, pts_pkg.LiDARPoints, pts_pkg.get_points_type = base.BasePoints, lidar.LiDARPoints, get_points_type
    sys.modules["mmdet3d.core"].points = pts_pkg

    class FileClient:
        def __init__(self, backend="disk", **_):
            assert backend == "disk"

        def get(self, path):
            with open(path, "rb") as f:
                return f.read()
    mmcv = sys.modules["mmcv"]
    mmcv.FileClient = FileClient
    mmcv.check_file_exist = lambda p: os.path.exists(p) or (_ for _ in ()).throw(FileNotFoundError(p))
    _pkg("mmdet.datasets")
    _mod("mmdet.datasets.builder", PIPELINES=Registry("pipeline"))
    _mod("mmdet.datasets.pipelines", LoadAnnotations=object, LoadImageFromFile=object)
    _pkg("mmdet3d.datasets")
    _pkg("mmdet3d.datasets.pipelines")
    env["loading"] = load("mmdet3d.datasets.pipelines.loading", "datasets/pipelines/loading.py")
    return env["loading"]


def load_finetune_consumer():
    """The reference's SSTInputLayer (middle_encoders/sst_input_layer.py) and SSTSecondPretrainedv1
    (backbones/sst_second_pretrained_v1.py) loaded by path on top of install(); mmcv's build_conv_layer is the one
    absent piece (restated: ``nn.Conv2d`` with the cfg's extra keys, which is all the 'Conv2d' type does).
    -> (SSTInputLayer class, SSTSecondPretrainedv1 class)."""
    env = install()
    if "finetune" in env:
        return env["finetune"]

    def build_conv_layer(cfg, *args, **kwargs):
        cfg = dict(cfg or dict(type="Conv2d"))
        assert cfg.pop("type") == "Conv2d"
        return nn.Conv2d(*args, **kwargs, **cfg)
    sys.modules["mmcv.cnn"].build_conv_layer = build_conv_layer

    def load(dotted, rel):
        spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF_ROOT, "mmdet3d", rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[dotted] = m
        spec.loader.exec_module(m)
        return m
    _pkg("mmdet3d.models.middle_encoders")
    inp = load("mmdet3d.models.middle_encoders.sst_input_layer", "models/middle_encoders/sst_input_layer.py")
    bb = load("mmdet3d.models.backbones.sst_second_pretrained_v1", "models/backbones/sst_second_pretrained_v1.py")
    env["finetune"] = (inp.SSTInputLayer, bb.SSTSecondPretrainedv1)
    return env["finetune"]


def load_config_model(config_rel: str = MAE_CONFIG) -> dict:
    """exec the reference config file and return its ``model`` dict."""
    ns: dict = {}
    with open(os.path.join(REF_ROOT, config_rel)) as f:
        exec(compile(f.read(), config_rel, "exec"), ns)
    return ns["model"]


def build_detector(model_cfg=None, seed=0, encoder_blocks=None, decoder_blocks=None):
    env = install()
    cfg = dict(model_cfg or load_config_model())
    if encoder_blocks is not None or decoder_blocks is not None:
        cfg["backbone"] = dict(cfg["backbone"])
        if encoder_blocks is not None:
            cfg["backbone"]["encoder_num_blocks"] = encoder_blocks
        if decoder_blocks is not None:
            cfg["backbone"]["decoder_num_blocks"] = decoder_blocks
    torch.manual_seed(seed)
    det = env["regs"]["DETECTORS"].build(cfg, default_args=dict(train_cfg=None, test_cfg=None))
    det.train()
    return det


class Recorder:
    """Wraps methods of a live reference detector to capture their outputs and to
    inject the keep/mask split (randperm differs between CPU and CUDA)."""

    def __init__(self, det, ids=None):
        self.det, self.out = det, {}
        self._ids = ids
        self._wrap(det, "voxelize")
        self._wrap(det, "sub_voxelize_low")
        self._wrap(det, "sub_voxelize_med")
        self._wrap(det, "voxel_encoder", call=True)
        self._wrap(det, "get_vanilla_mask_index", override=self._mask)
        self._wrap(det, "get_centroid_per_voxel", multi=True)
        self._wrap(det, "get_multi_voxel_id_to_tensor_id_for_curv")
        self._wrap(det, "cal_regular_voxel_nor_and_curv")
        self._wrap(det, "get_multi_voxel_id_to_tensor_id_ori")
        self._wrap(det, "extract_feat")
        self._wrap(det, "backbone", call=True)
        mod = sys.modules["mmdet3d.models.detectors.multi_sub_voxel_dynamic_voxelnet_ssl"]
        orig = mod.get_indice_pairs_implicit_gemm

        def pairs(*a, **k):
            r = orig(*a, **k)
            self.out["pair"] = r[2]
            return r
        mod.get_indice_pairs_implicit_gemm = pairs
        self._restore = lambda: setattr(mod, "get_indice_pairs_implicit_gemm", orig)

    def _mask(self, orig, coors, batch_size):
        if self._ids is not None:
            return self._ids
        return orig(coors, batch_size)

    def _wrap(self, obj, name, call=False, override=None, multi=False):
        target = getattr(obj, name)
        fn = target.forward if call else target

        def wrapped(*a, **k):
            r = override(fn, *a, **k) if override else fn(*a, **k)
            if multi:
                self.out.setdefault(name, []).append(r)
            else:
                self.out[name] = r
            return r
        if call:
            target.forward = wrapped
        else:
            setattr(obj, name, wrapped)

    def close(self):
        self._restore()


def vanilla_mask_ids(feature_coors, batch_size, ratio, seed):
    """The reference's per-sample random split (…_ssl.py:287-304) with a seeded CPU generator."""
    g = torch.Generator().manual_seed(seed)
    keep, mask = [], []
    for b in range(batch_size):
        inds = torch.where(feature_coors[:, 0] == b)[0]
        n = inds.shape[0]
        len_keep = int(n * (1 - ratio))
        perm = torch.randperm(n, generator=g)
        keep.append(inds[perm[:len_keep]])
        mask.append(inds[perm[len_keep:]])
    return torch.cat(keep), torch.cat(mask)
