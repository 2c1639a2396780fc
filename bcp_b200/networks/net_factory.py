"""Drop-in for the reference's ``networks/net_factory.py`` (/root/reference/code/networks/net_factory.py:5-19):
same signatures, same returned module kinds, already moved to the GPU."""
from .unet import UNet, UNet_2d
from .VNet import VNet


def net_factory(net_type="unet", in_chns=1, class_num=2, mode="train", tsne=0):
    net = None
    if net_type == "unet" and mode == "train":
        net = UNet(in_chns=in_chns, class_num=class_num).cuda()
    if net_type == "VNet" and mode == "train" and tsne == 0:
        net = VNet(n_channels=in_chns, n_classes=class_num, normalization="batchnorm", has_dropout=True).cuda()
    if net_type == "VNet" and mode == "test" and tsne == 0:
        net = VNet(n_channels=in_chns, n_classes=class_num, normalization="batchnorm", has_dropout=False).cuda()
    if net is None:
        raise ValueError("net_factory: unsupported (net_type=%r, mode=%r, tsne=%r)" % (net_type, mode, tsne))
    return net


def BCP_net(in_chns=1, class_num=2, ema=False):
    net = UNet_2d(in_chns=in_chns, class_num=class_num).cuda()
    if ema:
        for param in net.parameters():
            param.detach_()
    return net
