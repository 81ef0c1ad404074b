"""Drop-in for /root/reference/hamgnn/models/hamgnn_output.py: same class names, constructor and forward contract."""
from hamgnn_b200.hamgnn_output import HamGNNPlusPlusOut, HamLayer  # noqa: F401

__all__ = ["HamGNNPlusPlusOut", "HamLayer"]
