"""B200-native Pancreas V-Net with the module API of /root/reference/code/pancreas/Vnet.py:92-194
(flat block attributes, InstanceNorm3d default, ``forward(input, turnoff_drop=False)`` returning a list)."""
from __future__ import annotations

import torch.nn as nn

from .. import ops
from ..networks.runtime import NetRuntime
from ..networks.VNet import _Stage3d, _ENC, _DEC


class VNet(nn.Module):
    def __init__(self, n_channels=1, n_classes=2, n_filters=16, normalization="instancenorm", has_dropout=False):
        super().__init__()
        self.has_dropout = has_dropout
        for name, kind, st, ci, co in _ENC + _DEC[:-1]:
            setattr(self, name, _Stage3d(kind, st, n_channels if ci is None else ci * n_filters, co * n_filters, normalization))
        if has_dropout:
            self.dropout = nn.Dropout3d(p=0.5)
        mods = [_Stage3d("same", 1, n_filters, n_filters, normalization)]
        if has_dropout:
            mods.append(nn.Dropout3d(p=0.5))
        mods.append(nn.Conv3d(n_filters, n_classes, 1, padding=0))
        self.branchs = nn.ModuleList([nn.Sequential(*mods)])
        rt = NetRuntime(self, ("",))
        object.__setattr__(self, "_rt", rt)
        for m in self.modules():
            if isinstance(m, _Stage3d):
                m._bind(rt)
        rt.register_conv(self.branchs[0][-1], None)

    @property
    def runtime(self) -> NetRuntime:
        return self._rt

    def encoder(self, input, use_dropout):
        x1 = self.block_one(input)
        x2 = self.block_two(self.block_one_dw(x1))
        x3 = self.block_three(self.block_two_dw(x2))
        x4 = self.block_four(self.block_three_dw(self._rt.boundary(x3, "block_three_dw.conv.0.weight")))
        x4_dw = self.block_four_dw(x4)
        scale = None
        if use_dropout and self.dropout.training:
            scale = NetRuntime.channel_dropout_scale(self.dropout, x4_dw.shape[0], x4_dw.shape[1] * 8, x4_dw.device)
        return [x1, x2, x3, x4, self._rt.boundary(self.block_five(x4_dw, chan_scale=scale), "block_five_up.conv.0.weight")]

    def decoder(self, features):
        x1, x2, x3, x4, x5 = features
        u = self.block_five_up(x5, residual=x4)
        u = self.block_six_up(self.block_six(u), residual=x3)
        u = self.block_seven_up(self.block_seven(u), residual=x2)
        u = self.block_eight_up(self.block_eight(u), residual=x1)
        out = []
        for branch in self.branchs:
            scale = None
            if len(branch) == 3 and branch[1].training:      # the branch's own Dropout3d always runs (pancreas/Vnet.py:127)
                scale = NetRuntime.channel_dropout_scale(branch[1], u.shape[0], u.shape[1] * 8, u.device)
            x9 = branch[0](u, chan_scale=scale)
            out.append(ops.Head.apply(x9, branch[-1].weight, branch[-1].bias, False))
        return out

    def forward(self, input, turnoff_drop=False):
        rt = self._rt
        rt.prepare()
        rt.spg = input.shape[0]       # one reference forward call = one BatchNorm group (InstanceNorm uses spg 1 by itself)
        try:
            return self.decoder(self.encoder(input, self.has_dropout and not turnoff_drop))
        finally:
            rt.spg = None
