"""Drop-in for /root/reference/hamgnn/models/hamgnn_conv.py: same class name, constructor and forward contract."""
from hamgnn_b200.hamgnn_conv import HamGNNConvE3  # noqa: F401

__all__ = ["HamGNNConvE3"]
