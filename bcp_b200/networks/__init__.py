from .net_factory import net_factory, BCP_net  # noqa: F401
from .VNet import VNet  # noqa: F401
from .unet import UNet, UNet_2d  # noqa: F401
