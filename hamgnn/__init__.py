"""Import-path shim: `hamgnn.models.hamgnn_conv`, `hamgnn.models.hamgnn_output` and `hamgnn.data.graph_data`
resolve to the B200 implementation, so the reference's `hamgnn/main.py:27-34` import block

    from .data.graph_data import graph_data_module
    from .models.hamgnn_conv import HamGNNConvE3
    from .models.hamgnn_output import HamGNNPlusPlusOut

works unchanged when this directory is placed on the path instead of (or in front of) the reference's `hamgnn/models`
and `hamgnn/data`.  Nothing else of the reference package is re-implemented here: the Lightning
harness, config parsing and the DFT tools keep living in the reference tree (out of scope).
"""
