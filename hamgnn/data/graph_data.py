"""Drop-in for the NPZ half of /root/reference/hamgnn/data/graph_data.py (`NPZGraphDataset` :96-185,
`graph_data_module` :187-522).  See hamgnn_b200/datamodule.py."""
from hamgnn_b200.datamodule import NPZGraphDataset, graph_data_module  # noqa: F401

__all__ = ["NPZGraphDataset", "graph_data_module"]
