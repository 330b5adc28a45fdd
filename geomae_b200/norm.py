"""naiveSyncBN1d — mirror of mmdet3d/ops/norm.py:28-86 (same registry key, same semantics).

Single process / eval: plain BatchNorm1d statistics.  Distributed training: every rank
contributes [mean, mean-of-squares] with EQUAL weight (one all_gather forward, one all_reduce
backward), var = E[x^2] - E[x]^2 — including the reference's quirk of not weighting ranks
by their point counts."""
import torch
from torch import distributed as dist
from torch import nn
from torch.autograd.function import Function

from .registry import NORM_LAYERS


class _AllReduceSum(Function):
    @staticmethod
    def forward(ctx, x):
        gathered = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
        dist.all_gather(gathered, x)
        return torch.stack(gathered, dim=0).sum(dim=0)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g)
        return g


@NORM_LAYERS.register_module("naiveSyncBN1d")
class NaiveSyncBatchNorm1d(nn.BatchNorm1d):
    def forward(self, x):
        if x.dtype != torch.float32:
            raise RuntimeError(f"input should be in float32 type, got {x.dtype}")
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1 or not self.training:
            return super().forward(x)
        if x.shape[0] == 0:
            raise RuntimeError("SyncBN does not support empty inputs")
        c = x.shape[1]
        mean = x.mean(dim=0)
        meansqr = (x * x).mean(dim=0)
        vec = _AllReduceSum.apply(torch.cat([mean, meansqr])) * (1.0 / dist.get_world_size())
        mean, meansqr = torch.split(vec, c)
        var = meansqr - mean * mean
        with torch.no_grad():
            self.running_mean += self.momentum * (mean.detach() - self.running_mean)
            self.running_var += self.momentum * (var.detach() - self.running_var)
        scale = self.weight * torch.rsqrt(var + self.eps)
        return x * scale.view(1, -1) + (self.bias - mean * scale).view(1, -1)


NORM_LAYERS.register_module("BN1d")(nn.BatchNorm1d)
NORM_LAYERS.register_module("BN")(nn.BatchNorm1d)


@NORM_LAYERS.register_module("naiveSyncBN2d")
class NaiveSyncBatchNorm2d(nn.BatchNorm2d):
    """mmdet3d/ops/norm.py:89-144 — the 4-D twin of naiveSyncBN1d (statistics over N, H, W; equal rank weights)."""

    def forward(self, x):
        if x.dtype != torch.float32:
            raise RuntimeError(f"input should be in float32 type, got {x.dtype}")
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1 or not self.training:
            return super().forward(x)
        if x.shape[0] == 0:
            raise RuntimeError("SyncBN does not support empty inputs")
        c = x.shape[1]
        mean = x.mean(dim=[0, 2, 3])
        meansqr = (x * x).mean(dim=[0, 2, 3])
        vec = _AllReduceSum.apply(torch.cat([mean, meansqr])) * (1.0 / dist.get_world_size())
        mean, meansqr = torch.split(vec, c)
        var = meansqr - mean * mean
        with torch.no_grad():
            self.running_mean += self.momentum * (mean.detach() - self.running_mean)
            self.running_var += self.momentum * (var.detach() - self.running_var)
        scale = self.weight * torch.rsqrt(var + self.eps)
        return x * scale.view(1, -1, 1, 1) + (self.bias - mean * scale).view(1, -1, 1, 1)


NORM_LAYERS.register_module("BN2d")(nn.BatchNorm2d)
