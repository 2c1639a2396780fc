"""B200-native V-Net with the module API of the reference's ``networks/VNet.py``.

Same constructor arguments, ``forward`` return tuple, parameter order and ``state_dict`` keys as
/root/reference/code/networks/VNet.py:241-290 (so the shipped ``models/LA/*.pth`` load, and
``update_ema_variables``'s zip over ``parameters()`` pairs the same tensors) -- but the forward pass
runs on hand-written sm_100a kernels over channel-blocked bf16 activations:

  conv (tcgen05/TMA implicit GEMM, or the CUDA-core kernels for Cin=1 / stride-2 / head)
    -> train-mode BatchNorm statistics (fixed-order two-stage reduce)
    -> fused normalise + ReLU (+ Dropout3d channel scale) (+ skip add)

The child modules (nn.Conv3d, nn.BatchNorm3d, ...) exist as *parameter holders* so initialisation
and key names are identical to the reference; their own ``forward`` is never called.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .runtime import NetRuntime

_NORM_CTOR = {
    "batchnorm": lambda c: nn.BatchNorm3d(c),
    "groupnorm": lambda c: nn.GroupNorm(num_groups=16, num_channels=c),
    "instancenorm": lambda c: nn.InstanceNorm3d(c),
}


class _Stage3d(nn.Module):
    """n_stages x [conv, (norm), ReLU] held in ``self.conv`` (an nn.Sequential) like the reference blocks
    ConvBlock / DownsamplingConvBlock / UpsamplingDeconvBlock (networks/VNet.py:6-119)."""

    def __init__(self, kind, n_stages, cin, cout, normalization):
        super().__init__()
        layers = []
        self.kind, self.normalization = kind, normalization
        for i in range(n_stages):
            c_in = cin if i == 0 else cout
            if kind == "same":
                layers.append(nn.Conv3d(c_in, cout, 3, padding=1))
            elif kind == "down":
                layers.append(nn.Conv3d(c_in, cout, 2, padding=0, stride=2))
            else:
                layers.append(nn.ConvTranspose3d(c_in, cout, 2, padding=0, stride=2))
            if normalization != "none":
                if normalization not in _NORM_CTOR:
                    raise AssertionError(normalization)
                layers.append(_NORM_CTOR[normalization](cout))
            layers.append(nn.ReLU(inplace=True))
        self.conv = nn.Sequential(*layers)
        self._per = 3 if normalization != "none" else 2
        self._n = n_stages
        self._rt = None

    def _bind(self, rt: NetRuntime):
        object.__setattr__(self, "_rt", rt)
        for i in range(self._n):
            conv = self.conv[i * self._per]
            if self.kind == "same":
                rt.register_conv(conv, None if conv.in_channels % 8 else (0, 1))
            else:
                rt.register_conv(conv, (0, 2, 3))

    def _stats_req(self, norm, n):
        """Description of the batch-statistic normalisation that follows a conv, for the fused-statistics epilogue."""
        if norm is None or isinstance(norm, nn.GroupNorm):
            return None
        if isinstance(norm, nn.BatchNorm3d):
            if not (norm.training or not norm.track_running_stats):
                return None
            return dict(gamma=norm.weight, beta=norm.bias, running_mean=norm.running_mean, running_var=norm.running_var,
                        nbt=norm.num_batches_tracked, spg=self._rt.spg or n, eps=norm.eps,
                        momentum=norm.momentum if norm.momentum is not None else 0.1)
        return dict(gamma=norm.weight, beta=norm.bias, running_mean=None, running_var=None, nbt=None, spg=1, eps=norm.eps,
                    momentum=0.0)

    def _norm_act(self, y, norm, chan_scale=None, residual=None, precomputed=None):
        rt = self._rt
        n = y.shape[0]
        if norm is None:
            return ops.NormAct.apply(y, None, None, None, None, None, "none", n, 0.0, 0.0, 0.0, chan_scale, None, 1.0, residual)
        if isinstance(norm, nn.GroupNorm):
            raise NotImplementedError("normalization='groupnorm' is constructible in the reference but used by no entry "
                                      "point; not implemented in the sm_100a path")
        if isinstance(norm, nn.BatchNorm3d):
            if norm.training or not norm.track_running_stats:
                spg = rt.spg or n
                mom = norm.momentum if norm.momentum is not None else 0.1
                return ops.NormAct.apply(y, norm.weight, norm.bias, norm.running_mean, norm.running_var,
                                         norm.num_batches_tracked, "batch", spg, norm.eps, mom, 0.0, chan_scale, None, 1.0, residual,
                                         precomputed)
            return ops.NormAct.apply(y, norm.weight, norm.bias, norm.running_mean, norm.running_var, None, "eval", n,
                                     norm.eps, 0.0, 0.0, chan_scale, None, 1.0, residual)
        # InstanceNorm3d(affine=False, track_running_stats=False): per-sample statistics in train and eval
        return ops.NormAct.apply(y, norm.weight, norm.bias, None, None, None, "batch", 1, norm.eps, 0.0, 0.0, chan_scale,
                                 None, 1.0, residual, precomputed)

    def forward(self, a, chan_scale=None, residual=None):
        """a: CB8 activation, or the planar fp32 network input for the first block.
        chan_scale (Dropout3d) and residual (skip add) apply to the LAST stage's output."""
        rt = self._rt
        for i in range(self._n):
            conv = self.conv[i * self._per]
            norm = self.conv[i * self._per + 1] if self._per == 3 else None
            last = i == self._n - 1
            req = None
            # a conv that feeds batch-statistic normalisation has an identically-zero bias gradient (ops._bias_grad)
            bz = norm is not None and not isinstance(norm, nn.GroupNorm) and (
                not isinstance(norm, nn.BatchNorm3d) or norm.training or not norm.track_running_stats)
            if self.kind == "same":
                if a.dtype != torch.bfloat16:            # network input, planar fp32
                    if conv.in_channels == 1:
                        y = ops.ConvFirst.apply(a, conv.weight, conv.bias, bz)
                    else:
                        raise NotImplementedError("first layer with n_channels != 1")
                else:
                    req = self._stats_req(norm, a.shape[0])
                    y = ops.ConvSame.apply(a, conv.weight, conv.bias, rt.pack(conv), (3, 3, 3), bz, req)
            elif self.kind == "down":
                y = ops.ConvDown2.apply(a, conv.weight, conv.bias, rt.pack(conv), bz)
            else:
                y = ops.ConvUp2.apply(a, conv.weight, conv.bias, rt.pack(conv), bz)
            a = self._norm_act(y, norm, chan_scale if last else None, residual if last else None,
                               req.get("out") if req else None)
        return a


_ENC = [("block_one", "same", 1, None, 1), ("block_one_dw", "down", 1, 1, 2),
        ("block_two", "same", 2, 2, 2), ("block_two_dw", "down", 1, 2, 4),
        ("block_three", "same", 3, 4, 4), ("block_three_dw", "down", 1, 4, 8),
        ("block_four", "same", 3, 8, 8), ("block_four_dw", "down", 1, 8, 16),
        ("block_five", "same", 3, 16, 16)]
_DEC = [("block_five_up", "up", 1, 16, 8), ("block_six", "same", 3, 8, 8),
        ("block_six_up", "up", 1, 8, 4), ("block_seven", "same", 3, 4, 4),
        ("block_seven_up", "up", 1, 4, 2), ("block_eight", "same", 2, 2, 2),
        ("block_eight_up", "up", 1, 2, 1), ("block_nine", "same", 1, 1, 1)]


class Encoder(nn.Module):
    """networks/VNet.py:145-186."""

    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization="none", has_dropout=False, has_residual=False):
        super().__init__()
        if has_residual:
            raise NotImplementedError("has_residual=True is used by no entry point")
        self.has_dropout = has_dropout
        for name, kind, st, ci, co in _ENC:
            setattr(self, name, _Stage3d(kind, st, n_channels if ci is None else ci * n_filters, co * n_filters, normalization))
        self.dropout = nn.Dropout3d(p=0.5, inplace=False)

    def forward(self, input):
        x1 = self.block_one(input)
        x2 = self.block_two(self.block_one_dw(x1))
        x3 = self.block_three(self.block_two_dw(x2))
        # data-parallel bucket boundary: when backward reaches this point the deep encoder layers (73 % of the parameters)
        # have final gradients and their all-reduce overlaps the backward of the three full-resolution blocks
        x4 = self.block_four(self.block_three_dw(self.block_three._rt.boundary(x3, "encoder.block_three_dw.conv.0.weight")))
        x4_dw = self.block_four_dw(x4)
        scale = None
        if self.has_dropout and self.dropout.training:
            n, c = x4_dw.shape[0], x4_dw.shape[1] * 8
            scale = NetRuntime.channel_dropout_scale(self.dropout, n, c, x4_dw.device, self.block_five._rt.spg)
        x5 = self.block_five(x4_dw, chan_scale=scale)
        return [x1, x2, x3, x4, x5]


class Decoder(nn.Module):
    """networks/VNet.py:189-239."""

    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization="none", has_dropout=False, has_residual=False):
        super().__init__()
        self.has_dropout = has_dropout
        for name, kind, st, ci, co in _DEC:
            setattr(self, name, _Stage3d(kind, st, ci * n_filters, co * n_filters, normalization))
        self.out_conv = nn.Conv3d(n_filters, n_classes, 1, padding=0)
        self.dropout = nn.Dropout3d(p=0.5, inplace=False)

    def forward(self, features):
        x1, x2, x3, x4, x5 = features
        x5_up = self.block_five_up(x5, residual=x4)
        x6_up = self.block_six_up(self.block_six(x5_up), residual=x3)
        x7_up = self.block_seven_up(self.block_seven(x6_up), residual=x2)
        x8_up = self.block_eight_up(self.block_eight(x7_up), residual=x1)
        scale = None
        if self.has_dropout and self.dropout.training:
            n, c = x8_up.shape[0], x8_up.shape[1] * 8
            scale = NetRuntime.channel_dropout_scale(self.dropout, n, c, x8_up.device, self.block_nine._rt.spg)
        x9 = self.block_nine(x8_up, chan_scale=scale)
        out_seg = ops.Head.apply(x9, self.out_conv.weight, self.out_conv.bias, False)
        return out_seg, x8_up


def _mlp_heads(owner, n_sel):
    """Never-trained heads kept for key/parameter-order parity (networks/VNet.py:250-278)."""
    owner.projection_head = nn.Sequential(nn.Linear(16, 32), nn.BatchNorm1d(32), nn.ReLU(inplace=True), nn.Linear(32, 32))
    owner.prediction_head = nn.Sequential(nn.Linear(32, 32), nn.BatchNorm1d(32), nn.ReLU(inplace=True), nn.Linear(32, 32))
    for stem in ("contrastive_class_selector_", "contrastive_class_selector_memory"):
        for c in range(n_sel):
            setattr(owner, stem + str(c), nn.Sequential(nn.Linear(32, 32), nn.BatchNorm1d(32),
                                                        nn.LeakyReLU(negative_slope=0.2, inplace=True), nn.Linear(32, 1)))


class VNet(nn.Module):
    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization="none", has_dropout=False, has_residual=False):
        super().__init__()
        self.encoder = Encoder(n_channels, n_classes, n_filters, normalization, has_dropout, has_residual)
        self.decoder = Decoder(n_channels, n_classes, n_filters, normalization, has_dropout, has_residual)
        self.pool = nn.MaxPool3d(3, stride=2)
        _mlp_heads(self, 2)
        rt = NetRuntime(self, ("encoder.", "decoder."))
        object.__setattr__(self, "_rt", rt)
        for m in self.modules():
            if isinstance(m, _Stage3d):
                m._bind(rt)
        rt.register_conv(self.decoder.out_conv, None)

    @property
    def runtime(self) -> NetRuntime:
        return self._rt

    def forward_projection_head(self, features):
        return self.projection_head(features)

    def forward_prediction_head(self, features):
        return self.prediction_head(features)

    def forward(self, input, groups: int = 1, with_features: bool = True):
        """input [N,1,X,Y,Z] fp32 -> (logits [N,n_classes,X,Y,Z] fp32, features).
        ``groups`` > 1 batches that many reference forward calls (BatchNorm statistics stay per call)."""
        rt = self._rt
        rt.prepare()
        n = input.shape[0]
        assert n % groups == 0
        rt.spg = n // groups
        try:
            feats = self.encoder(input)
            feats[4] = rt.boundary(feats[4], "decoder.block_five_up.conv.0.weight")      # bucket boundary: decoder done
            out_seg, _ = self.decoder(feats)
            x5 = feats[4]
            features = None
            if with_features and min(x5.shape[2:5]) >= 3:
                features = ops.maxpool3d_k3s2(x5, x5.shape[1] * 8)
        finally:
            rt.spg = None
        return out_seg, features
