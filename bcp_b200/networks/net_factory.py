"""Constructors the entry scripts call (API of the reference's networks/net_factory.py:5-19): ``net_factory`` for the LA
V-Net / 2-D U-Net and ``BCP_net`` for the ACDC student/teacher pair.  Modules come back on the current CUDA device."""
from . import unet as _unet
from . import VNet as _vnet

# (net_type, mode) -> builder(in_chns, class_num); the t-SNE variant of the reference (tsne != 0) is not part of the hot path
_BUILDERS = {
    ("unet", "train"): lambda c, k: _unet.UNet(in_chns=c, class_num=k),
    ("VNet", "train"): lambda c, k: _vnet.VNet(n_channels=c, n_classes=k, normalization="batchnorm", has_dropout=True),
    ("VNet", "test"): lambda c, k: _vnet.VNet(n_channels=c, n_classes=k, normalization="batchnorm", has_dropout=False),
}


def net_factory(net_type="unet", in_chns=1, class_num=2, mode="train", tsne=0):
    key = (net_type, mode)
    if key not in _BUILDERS or (net_type == "VNet" and tsne != 0):
        raise ValueError("net_factory: unsupported (net_type=%r, mode=%r, tsne=%r)" % (net_type, mode, tsne))
    return _BUILDERS[key](in_chns, class_num).cuda()


def BCP_net(in_chns=1, class_num=2, ema=False):
    """2-D U-Net for ACDC; ``ema=True`` returns the teacher, whose parameters take no gradients."""
    model = _unet.UNet_2d(in_chns=in_chns, class_num=class_num).cuda()
    if ema:
        for weight in model.parameters():
            weight.detach_()
    return model
