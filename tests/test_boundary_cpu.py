"""Drop-in boundary on the CPU: the `hamgnn.*` import path of the reference's main.py (hamgnn/main.py:27-34), module
construction from the reference's default config (hamgnn/config/config_parsing.py:46-91) and from the Uni-HamGNN
driver's keyword set (Uni-HamGNN/Uni-HamiltonianPredictor.py:40-76), the graph_data.npz formats
(DFT_interfaces/openmx/graph_data_gen.py:357-380; loader hamgnn/data/graph_data.py:110-159) and the
get_nonzero_mask_tensor outputs (hamgnn_output.py:2588-2782).  No kernel runs here."""
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from hamgnn_b200 import graph_data as gd


class _NS(dict):
    """EasyDict-like config node (attribute + item access), as main.py passes it."""
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def _ns(d):
    return _NS({k: _ns(v) if isinstance(v, dict) else v for k, v in d.items()})


# hamgnn/config/config_parsing.py:46-91
PRE_DEFAULTS = dict(cutoff=26.0, cutoff_func="cos", radius_type="openmx", edge_sh_normalization="component",
                    edge_sh_normalize=True, irreps_edge_sh="0e + 1o + 2e + 3o + 4e + 5o",
                    irreps_node_features="64x0e+64x0o+32x1o+16x1e+12x2o+25x2e+18x3o+9x3e+4x4o+9x4e+4x5o+4x5e+2x6e",
                    num_layers=3, num_radial=64, num_types=96, rbf_func="bessel", set_features=True, radial_MLP=[64, 64],
                    use_corr_prod=False, correlation=2, num_hidden_features=16, use_kan=False, radius_scale=1.01,
                    build_internal_graph=False, use_gradient_checkpointing=False)
OUT_DEFAULTS = dict(ham_only=True, ham_type="openmx", nao_max=26, add_H0=True, add_H_nonsoc=False, symmetrize=True,
                    calculate_band_energy=False, num_k=5, band_num_control=8, k_path=None, soc_switch=False,
                    nonlinearity_type="gate", spin_constrained=False, collinear_spin=False, minMagneticMoment=0.5,
                    zero_point_shift=True, get_nonzero_mask_tensor=False)


def test_reference_import_block_and_default_config():
    from hamgnn.data.graph_data import graph_data_module  # noqa: F401  (main.py:27)
    from hamgnn.models.hamgnn_conv import HamGNNConvE3    # main.py:32
    from hamgnn.models.hamgnn_output import HamGNNPlusPlusOut  # main.py:33
    cfg = _ns({"representation_nets": {"HamGNN_pre": dict(PRE_DEFAULTS)}, "output_nets": {"HamGNN_out": dict(OUT_DEFAULTS)}})
    cfg.representation_nets.HamGNN_pre.radius_type = cfg.output_nets.HamGNN_out.ham_type.lower()   # main.py:211
    pre = HamGNNConvE3(cfg.representation_nets)
    o = cfg.output_nets.HamGNN_out
    out = HamGNNPlusPlusOut(irreps_in_node=pre.irreps_node_features, irreps_in_edge=pre.irreps_node_features,
                            nao_max=o.nao_max, ham_type=o.ham_type, ham_only=o.ham_only, symmetrize=o.symmetrize,
                            calculate_band_energy=o.calculate_band_energy, num_k=o.num_k, k_path=o.k_path,
                            band_num_control=o.band_num_control, soc_switch=o.soc_switch, soc_basis="so3",
                            nonlinearity_type=o.nonlinearity_type, add_H0=o.add_H0, spin_constrained=o.spin_constrained,
                            collinear_spin=o.collinear_spin, minMagneticMoment=o.minMagneticMoment,
                            add_H_nonsoc=o.add_H_nonsoc, get_nonzero_mask_tensor=o.get_nonzero_mask_tensor,
                            zero_point_shift=o.zero_point_shift)                                   # main.py:234-255
    assert str(pre.irreps_node_features).replace(" ", "") == PRE_DEFAULTS["irreps_node_features"]
    assert out.nao_max == 26 and out.derivative is False and out.zero_point_shift is True
    assert sum(p.numel() for p in pre.parameters()) > 1_000_000
    assert all(isinstance(p, torch.nn.Parameter) for p in out.parameters())


def test_uni_hamgnn_constructor_keywords():
    """Uni-HamiltonianPredictor.py:40-76: legacy_edge_update=True, use_corr_prod=False, nao 26, get_nonzero_mask_tensor=True."""
    from hamgnn.models.hamgnn_conv import HamGNNConvE3
    from hamgnn.models.hamgnn_output import HamGNNPlusPlusOut
    rep = _ns({"HamGNN_pre": dict(PRE_DEFAULTS, irreps_node_features="8x0e+8x0o+4x1o+4x1e+3x2o+5x2e+2x3o+2x3e+2x4e",
                                  num_layers=2, num_radial=16, radial_MLP=[16, 16], irreps_edge_sh="0e+1o+2e+3o+4e")})
    rep.HamGNN_pre.radius_type = "openmx"
    rep.HamGNN_pre.ham_type = "openmx"
    rep.HamGNN_pre.nao_max = 26
    rep.HamGNN_pre.use_corr_prod = False
    rep.HamGNN_pre.legacy_edge_update = True
    pre = HamGNNConvE3(rep)
    assert pre.legacy_edge_update and not hasattr(pre.pair_interactions[0], "skip_linear")
    for soc in (False, True):
        out = HamGNNPlusPlusOut(irreps_in_node=pre.irreps_node_features, irreps_in_edge=pre.irreps_node_features, nao_max=26,
                                ham_type="openmx", ham_only=True, symmetrize=True, calculate_band_energy=False, num_k=5,
                                k_path=None, band_num_control=8, soc_switch=soc, nonlinearity_type="gate", add_H0=True,
                                spin_constrained=False, collinear_spin=False, minMagneticMoment=0.5,
                                add_H_nonsoc=True if soc else False, zero_point_shift=False if soc else True,
                                get_nonzero_mask_tensor=True)
        assert out.get_nonzero_mask_tensor
        out.zero_point_shift = False   # the Uni driver mutates it (Uni-HamiltonianPredictor.py:248-251)


def _mask_oracle(basis_def, nao, z, src, dst):
    """Loop restatement of create_orbital_validity_mask + build_interaction_masks (hamgnn_output.py:2588-2665)."""
    tab = np.zeros((99, nao), dtype=bool)
    for Z, orbs in basis_def.items():
        tab[Z, list(orbs)] = True
    on = np.stack([np.outer(tab[a], tab[a]).reshape(-1) for a in z])
    off = np.stack([np.outer(tab[z[i]], tab[z[j]]).reshape(-1) for i, j in zip(src, dst)])
    return on, off


@pytest.mark.parametrize("nao", [19, 26])
def test_nonzero_masks_match_loop_restatement(nao):
    from hamgnn_b200.hamgnn_output import HamGNNPlusPlusOut
    g1 = gd.mos2_monolayer(seed=0, nao_max=nao)
    g2 = gd.bulk_silicon(seed=1, nao_max=nao)
    b = gd.Batch.from_data_list([g1, g2])
    D = "4x0e+4x1o"
    out = HamGNNPlusPlusOut(D, D, nao_max=nao, soc_switch=False, ham_only=True, get_nonzero_mask_tensor=True)
    z, (src, dst) = b.z.numpy(), b.edge_index.numpy()
    on, off = _mask_oracle(out.basis_def, nao, z, src, dst)
    m = out.build_interaction_masks(b)
    assert m.dtype == torch.bool and tuple(m.shape) == (len(z) + len(src), nao * nao)
    assert np.array_equal(m.numpy(), np.concatenate([on, off]))          # plain cat: on-site rows first (:2660-2665)
    mc = out.build_column_wise_interaction_masks(b)
    assert tuple(mc.shape) == (len(z) + len(src), 2, nao * nao) and torch.equal(mc[:, 0], m) and torch.equal(mc[:, 1], m)
    soc = HamGNNPlusPlusOut(D, D, nao_max=nao, soc_switch=True, soc_basis="so3", ham_only=True, get_nonzero_mask_tensor=True)
    ri, full = soc.build_spin_orbit_interaction_masks(b)
    M = 2 * nao
    assert tuple(ri.shape) == (len(z) + len(src), M * M) and tuple(full.shape) == (2 * (len(z) + len(src)), M * M)
    on4 = np.stack([np.tile(x.reshape(nao, nao), (2, 2)).reshape(-1) for x in on])
    off4 = np.stack([np.tile(x.reshape(nao, nao), (2, 2)).reshape(-1) for x in off])
    want = soc.concatenate_hamiltonians_by_crystal(b, torch.from_numpy(on4), torch.from_numpy(off4))   # per-crystal rows (:2775)
    assert torch.equal(ri, want) and torch.equal(full, torch.cat([want, want]))
    # masked entries of the symmetric synthetic targets are exactly the orbitals missing from basis_def
    assert int(m.sum()) == int(on.sum() + off.sum()) < m.numel()


# ------------------------------------------------------------------------------------------------ graph_data.npz
def _graphs():
    return [gd.bulk_silicon(seed=i) for i in range(5)]


def _same(a: gd.Data, b: gd.Data):
    assert sorted(a.keys()) == sorted(b.keys())
    for k in a.keys():
        va, vb = a[k], b[k]
        if torch.is_tensor(va):
            assert va.dtype == vb.dtype and torch.equal(va, vb), k
        else:
            assert va == vb, k


def test_npz_dict_flavour_roundtrip(tmp_path):
    path = str(tmp_path / "graph_data.npz")
    gs = _graphs()
    gd.save_graph_data_npz(path, gs)
    back = gd.load_graph_data_npz(path)
    assert len(back) == len(gs)
    for a, b in zip(gs, back):
        _same(a, b)
    batch = gd.Batch.from_data_list(back[:2])
    n0 = gs[0].num_nodes
    assert torch.equal(batch.edge_index[:, gs[0].edge_index.shape[1]:], gs[1].edge_index + n0)   # 'index' keys are offset
    assert torch.equal(batch.inv_edge_idx, torch.cat([gs[0].inv_edge_idx, gs[1].inv_edge_idx]))  # ... inv_edge_idx is not


def _fake_pyg_modules():
    """A torch_geometric >= 2.0 look-alike: Data.__dict__ = {'_store': GlobalStorage}, GlobalStorage.__dict__ =
    {'_mapping': {...}, '_parent': Data} -- the pickle layout graph_data_gen.py:357-380 leaves on disk."""
    mods = {n: types.ModuleType(n) for n in ("torch_geometric", "torch_geometric.data", "torch_geometric.data.data",
                                             "torch_geometric.data.storage")}

    class GlobalStorage:
        def __init__(self, mapping, parent):
            self._mapping, self._parent = dict(mapping), parent

    class Data:
        def __init__(self, **kw):
            self._edge_attr_cls, self._tensor_attr_cls = "EdgeAttr", "TensorAttr"
            self._store = GlobalStorage(kw, self)

    GlobalStorage.__module__, GlobalStorage.__qualname__ = "torch_geometric.data.storage", "GlobalStorage"
    Data.__module__, Data.__qualname__ = "torch_geometric.data.data", "Data"
    mods["torch_geometric.data.storage"].GlobalStorage = GlobalStorage
    mods["torch_geometric.data.data"].Data = Data
    mods["torch_geometric.data"].Data = Data
    return mods, Data


def test_npz_pickled_pyg_data_objects(tmp_path):
    path = str(tmp_path / "graph_data.npz")
    gs = _graphs()[:3]
    mods, FakeData = _fake_pyg_modules()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        np.savez(path, graph={i: FakeData(**g.to_dict()) for i, g in enumerate(gs)})
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    back = gd.load_graph_data_npz(path)     # no torch_geometric in this image: the loader's stub revives the objects
    assert len(back) == 3
    for a, b in zip(gs, back):
        _same(a, b)


def test_graph_data_module_npz(tmp_path):
    from hamgnn.data.graph_data import NPZGraphDataset, graph_data_module
    path = str(tmp_path / "graph_data.npz")
    gs = [gd.bulk_silicon(seed=i) for i in range(10)]
    gd.save_graph_data_npz(path, gs)
    split = str(tmp_path / "split.npz")
    dm = graph_data_module(dataset=path, train_ratio=0.6, val_ratio=0.2, test_ratio=0.2, batch_size=2, split_file=split,
                           num_workers=0)
    dm.prepare_data()
    dm.setup("fit")
    dm.setup("test")
    idx = list(range(10))
    np.random.RandomState(seed=42).shuffle(idx)                     # hamgnn/data/graph_data.py:367-380
    assert dm.train_data.indices == idx[:6] and dm.val_data.indices == idx[6:8] and dm.test_data.indices == idx[8:]
    sp = np.load(split)
    assert sp["train_idx"].tolist() == idx[:6]
    b = next(iter(dm.val_dataloader()))
    assert isinstance(b, gd.Batch) and b.num_graphs == 2 and b.z.shape[0] == 4
    _same(gd.Batch.from_data_list([gs[i] for i in idx[6:8]]), b)
    dm2 = graph_data_module(dataset=path, batch_size=4, split_file=split, num_workers=0)    # re-uses the saved split
    dm2.setup()
    assert dm2.test_data.indices == idx[8:]
    tm = graph_data_module(dataset=gs, test_mode=True, batch_size=3, num_workers=0)         # in-memory list, test mode
    tm.setup()
    assert len(tm.test_data) == 10 and len(tm.train_data) == 0
    assert sum(bb.num_graphs for bb in tm.test_dataloader()) == 10
    assert len(NPZGraphDataset(path, indices=[1, 3], preload=1)) == 2
    with pytest.raises(NotImplementedError):
        graph_data_module(dataset="x.lmdb").setup()
    with pytest.raises(ValueError):
        graph_data_module(dataset="x.bin").setup()


def test_strict_reference_state_dict_loader():
    """load_reference_state_dict: e3nn's constant buffers are accepted, a renamed / missing / mis-shaped weight raises
    (`strict=False` drops mismatched names silently)."""
    from hamgnn_b200.hamgnn_conv import HamGNNConvE3, load_reference_state_dict
    cfg = dict(irreps_node_features="8x0e+8x0o+4x1o+4x1e+3x2o+5x2e+2x3o+2x3e+2x4e", num_layers=2, num_radial=16, radial_MLP=[16, 16],
               irreps_edge_sh="0e+1o+2e+3o+4e")
    torch.manual_seed(1)
    src = HamGNNConvE3(cfg)
    sd = dict(src.state_dict())
    # what an e3nn 0.5.0 checkpoint carries in addition
    sd["convolutions.0.conv_tp.node_tensor_product.output_mask"] = torch.ones(10)
    sd["convolutions.0.conv_tp.node_tensor_product._compiled_main_left_right._w3j_1_1_2"] = torch.zeros(3, 3, 5)
    sd["convolutions.0.skip_linear.bias"] = torch.zeros(0)
    sd["pair_embedding.linear_up_src.output_mask"] = torch.ones(96)
    dst = HamGNNConvE3(cfg)
    load_reference_state_dict(dst, sd)
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    bad = dict(sd)
    bad["convolutions.0.conv_tp.node_tensor_product.weights"] = bad.pop("convolutions.0.conv_tp.node_tensor_product.weight")   # renamed
    with pytest.raises(RuntimeError):
        load_reference_state_dict(dst, bad)
    bad = dict(sd)
    bad["chemical_embedding.linear.weight"] = torch.zeros(3)
    with pytest.raises(RuntimeError):
        load_reference_state_dict(dst, bad)
    bad = dict(sd)
    bad["convolutions.0.skip_linear.bias"] = torch.zeros(4)       # a non-empty bias is not something the kernels evaluate
    with pytest.raises(RuntimeError):
        load_reference_state_dict(dst, bad)
