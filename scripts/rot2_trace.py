import os, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from hamgnn_b200 import graph_data as gd
from hamgnn_b200.hamgnn_conv import HamGNNConvE3
torch.manual_seed(0)
pre = HamGNNConvE3({}).cuda()
g = gd.Batch.from_data_list([gd.twisted_bilayer_graphene(m=4, seed=0)]).to('cuda')
with torch.no_grad():
    pre(g)            # warm
    torch.cuda.synchronize()
    os.environ["HGB_ROT2_TRACE"] = "gpurun_out/r02f_trace.txt"
    cb = pre.convolutions[0]
    g2 = gd.Batch(**g.to_dict())
    pre(g2)
    torch.cuda.synchronize()
