"""B200-native 2-D U-Net with the module API of the reference's ``networks/unet.py`` (UNet / UNet_2d).

Constructor arguments, forward return values, parameter order and state_dict keys follow
/root/reference/code/networks/unet.py:15-116,148-257 so ``models/ACDC/*.pth`` load.  The forward pass runs
on the sm_100a kernels over CB8 bf16 activations with X == 1 (a 3x3 conv is a 1x3x3 conv).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .runtime import NetRuntime
from .VNet import _mlp_heads


class ConvBlock(nn.Module):
    """conv-bn-lrelu-dropout-conv-bn-lrelu (networks/unet.py:15-30)."""

    def __init__(self, in_channels, out_channels, dropout_p):
        super().__init__()
        self.conv_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels), nn.LeakyReLU(),
            nn.Dropout(dropout_p),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels), nn.LeakyReLU())
        self._rt = None

    def _bind(self, rt):
        object.__setattr__(self, "_rt", rt)
        for i in (0, 4):
            conv = self.conv_conv[i]
            rt.register_conv(conv, None if conv.in_channels % 8 else (0, 1))

    def _stats_req(self, bn, n):
        """the batch-statistic BatchNorm that follows a conv, for the fused-statistics conv epilogue (ops.ConvSame)"""
        if not (bn.training or not bn.track_running_stats):
            return None
        return dict(gamma=bn.weight, beta=bn.bias, running_mean=bn.running_mean, running_var=bn.running_var,
                    nbt=bn.num_batches_tracked, spg=self._rt.spg or n, eps=bn.eps,
                    momentum=bn.momentum if bn.momentum is not None else 0.1)

    def _bn_act(self, y, bn, slope, keep=None, keep_scale=1.0, precomputed=None):
        rt = self._rt
        n = y.shape[0]
        if bn.training or not bn.track_running_stats:
            mom = bn.momentum if bn.momentum is not None else 0.1
            return ops.NormAct.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, "batch",
                                     rt.spg or n, bn.eps, mom, slope, None, keep, keep_scale, None, precomputed)
        return ops.NormAct.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, None, "eval", n, bn.eps, 0.0, slope,
                                 None, keep, keep_scale, None)

    def forward(self, a):
        rt = self._rt
        c0, bn0, act0, drop, c1, bn1, act1 = self.conv_conv
        bz0 = bn0.training or not bn0.track_running_stats       # batch-stat BN next: bias gradient is identically zero
        bz1 = bn1.training or not bn1.track_running_stats
        req0 = None
        if a.dtype != torch.bfloat16:
            if c0.in_channels != 1:
                raise NotImplementedError("first layer with in_chns != 1")
            y = ops.ConvFirst.apply(a, c0.weight, c0.bias, bz0)
        else:
            req0 = self._stats_req(bn0, a.shape[0])
            y = ops.ConvSame.apply(a, c0.weight, c0.bias, rt.pack(c0), (1, 3, 3), bz0, req0)
        keep, scale = None, 1.0
        if drop.training:
            n, cb, x, yy, z, _ = y.shape
            keep, scale = NetRuntime.element_dropout_keep(drop, n, cb * 8, (x, yy, z), y.device, rt.spg)
        a = self._bn_act(y, bn0, act0.negative_slope, keep, scale, req0.get("out") if req0 else None)
        req1 = self._stats_req(bn1, a.shape[0])
        y = ops.ConvSame.apply(a, c1.weight, c1.bias, rt.pack(c1), (1, 3, 3), bz1, req1)
        return self._bn_act(y, bn1, act1.negative_slope, precomputed=req1.get("out") if req1 else None)


class DownBlock(nn.Module):
    """MaxPool2d(2) then ConvBlock (networks/unet.py:32-43)."""

    def __init__(self, in_channels, out_channels, dropout_p):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), ConvBlock(in_channels, out_channels, dropout_p))

    def forward(self, a):
        return self.maxpool_conv[1](ops.MaxPool2.apply(a))


class UpBlock(nn.Module):
    """1x1 conv, bilinear x2 (align_corners=True), concat with the skip, ConvBlock (networks/unet.py:45-57)."""

    def __init__(self, in_channels1, in_channels2, out_channels, dropout_p):
        super().__init__()
        self.conv1x1 = nn.Conv2d(in_channels1, in_channels2, kernel_size=1)
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = ConvBlock(in_channels2 * 2, out_channels, dropout_p)
        self._rt = None

    def _bind(self, rt):
        object.__setattr__(self, "_rt", rt)
        rt.register_conv(self.conv1x1, (0, 1))

    def forward(self, x1, x2):
        rt = self._rt
        x1 = ops.ConvSame.apply(x1, self.conv1x1.weight, self.conv1x1.bias, rt.pack(self.conv1x1), (1, 1, 1))
        x1 = ops.Upsample2.apply(x1)
        x = torch.cat([x2, x1], dim=1)          # CB8: channel blocks are dim 1, so this is the reference's channel concat
        return self.conv(x)


class Encoder(nn.Module):
    def __init__(self, params):
        super().__init__()
        self.params = params
        self.in_chns, self.ft_chns = params["in_chns"], params["feature_chns"]
        self.n_class, self.dropout = params["class_num"], params["dropout"]
        assert len(self.ft_chns) == 5
        self.in_conv = ConvBlock(self.in_chns, self.ft_chns[0], self.dropout[0])
        self.down1 = DownBlock(self.ft_chns[0], self.ft_chns[1], self.dropout[1])
        self.down2 = DownBlock(self.ft_chns[1], self.ft_chns[2], self.dropout[2])
        self.down3 = DownBlock(self.ft_chns[2], self.ft_chns[3], self.dropout[3])
        self.down4 = DownBlock(self.ft_chns[3], self.ft_chns[4], self.dropout[4])

    def forward(self, x):
        x0 = self.in_conv(x)
        x1 = self.down1(x0)
        x2 = self.down2(x1)
        x3 = self.down3(x2)
        x4 = self.down4(x3)
        return [x0, x1, x2, x3, x4]


class Decoder(nn.Module):
    def __init__(self, params):
        super().__init__()
        self.params = params
        self.in_chns, self.ft_chns, self.n_class = params["in_chns"], params["feature_chns"], params["class_num"]
        assert len(self.ft_chns) == 5
        f = self.ft_chns
        self.up1 = UpBlock(f[4], f[3], f[3], dropout_p=0.0)
        self.up2 = UpBlock(f[3], f[2], f[2], dropout_p=0.0)
        self.up3 = UpBlock(f[2], f[1], f[1], dropout_p=0.0)
        self.up4 = UpBlock(f[1], f[0], f[0], dropout_p=0.0)
        self.out_conv = nn.Conv2d(f[0], self.n_class, kernel_size=3, padding=1)

    def forward(self, feature):
        x0, x1, x2, x3, x4 = feature
        x = self.up1(x4, x3)
        x = self.up2(x, x2)
        x = self.up3(x, x1)
        x_last = self.up4(x, x0)
        output = ops.Head.apply(x_last, self.out_conv.weight, self.out_conv.bias, True)
        return output, x_last


class _UNetBase(nn.Module):
    def __init__(self, in_chns, class_num):
        super().__init__()
        params = {"in_chns": in_chns, "feature_chns": [16, 32, 64, 128, 256], "dropout": [0.05, 0.1, 0.2, 0.3, 0.5],
                  "class_num": class_num, "acti_func": "relu"}
        self.encoder = Encoder(params)
        self.decoder = Decoder(params)
        _mlp_heads(self, 4)
        rt = NetRuntime(self, ("encoder.", "decoder."))
        object.__setattr__(self, "_rt", rt)
        for m in self.modules():
            if isinstance(m, (ConvBlock, UpBlock)):
                m._bind(rt)
        rt.register_conv(self.decoder.out_conv, None)

    @property
    def runtime(self) -> NetRuntime:
        return self._rt

    def forward_projection_head(self, features):
        return self.projection_head(features)

    def forward_prediction_head(self, features):
        return self.prediction_head(features)

    def _run(self, x, groups):
        rt = self._rt
        rt.prepare()
        n = x.shape[0]
        assert n % groups == 0
        rt.spg = n // groups
        try:
            return self.decoder(self.encoder(x))
        finally:
            rt.spg = None


class UNet(_UNetBase):
    """networks/unet.py:148-201: returns (logits, x_last)."""

    def forward(self, x, groups: int = 1):
        output, x_last = self._run(x, groups)
        return output, ops.CB8ToPlanar.apply(x_last, x_last.shape[1] * 8, True)


class UNet_2d(_UNetBase):
    """networks/unet.py:203-257: returns logits only."""

    def forward(self, x, groups: int = 1):
        return self._run(x, groups)[0]
