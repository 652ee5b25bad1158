from masic_b200.layers import GDN, MaskedConv2d, ResidualBlock, conv1x1, conv3x3  # noqa: F401

__all__ = ["GDN", "MaskedConv2d", "ResidualBlock", "conv3x3"]
