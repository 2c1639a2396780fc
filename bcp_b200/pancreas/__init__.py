from .Vnet import VNet  # noqa: F401
