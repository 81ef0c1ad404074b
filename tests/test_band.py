"""Band-energy head (hamgnn/models/hamgnn_output.py:1675-1996).

CPU: the oracle restatement (oracle/band_ref.py) against independent closed-form checks -- with S = 1 the bands are the
eigenvalues of the Hermitian Bloch sum, they do not depend on the order of the edges, and a one-orbital nearest-neighbour chain
gives e(k) = e0 + 2 t cos(2 pi k).  GPU: hgb_band_kspace + the cuSOLVER path of hamgnn_b200.band against the oracle on a two-crystal
batch with the model's own (Hermitian) Hamiltonian blocks and a synthetic positive-definite overlap, and through
HamGNNPlusPlusOut(calculate_band_energy=True).  Tolerances: H(k), S(k) 2e-6 relative (fp32 sums of <= 30 images vs fp64),
band energies / gap 2e-4 of the spectral width (fp32 Cholesky + eigh of a 100 x 100 pencil vs fp64)."""
import numpy as np
import pytest
import torch

from hamgnn_b200 import graph_data as gd
from oracle import band_ref


def _global_inverse(batch):
    """inv_edge_idx is per graph (reference collation, graph_data_gen.py:293-295): add every crystal's edge offset."""
    eb = batch.batch[batch.edge_index[0]]
    counts = torch.bincount(eb)
    off = torch.cumsum(counts, 0) - counts
    return batch.inv_edge_idx + off[eb]


def _pd_overlap(batch, nao, seed=0, eps=0.01):
    """Synthetic overlap blocks: Son = 1 + small symmetric, Soff[e] = eps * A with Soff[inv e] = Soff[e]^T (S(k) Hermitian, PD)."""
    g = torch.Generator().manual_seed(seed)
    N, E = batch.num_nodes, batch.edge_index.shape[1]
    a = torch.randn(N, nao, nao, generator=g) * eps
    son = torch.eye(nao).expand(N, nao, nao) + 0.5 * (a + a.transpose(1, 2))
    b = torch.randn(E, nao, nao, generator=g) * eps
    inv = _global_inverse(batch)
    soff = 0.5 * (b + b[inv].transpose(1, 2))
    return son.reshape(N, -1).contiguous(), soff.reshape(E, -1).contiguous()


def test_oracle_chain_dispersion():
    # one atom with two decoupled orbitals, neighbours at +-a: e_0(k) = e0 + 2 t cos(2 pi k a), e_1 = 10 (flat); gap between them
    e0, t = 0.3, -1.1
    ks = np.linspace(-0.5, 0.5, 9)
    kv = np.stack([ks, 0 * ks, 0 * ks], 1)[None]
    hon = np.array([[e0, 0, 0, 10.0]])
    hoff = np.array([[t, 0, 0, 0.0], [t, 0, 0, 0.0]])
    out = band_ref.band_energies(hon, hoff, np.array([[1.0, 0, 0, 1.0]]), np.zeros((2, 4)), np.array([[0, 0], [0, 0]]),
                                 np.array([1]), np.array([0]), [1], np.array([[1.0, 0, 0], [-1.0, 0, 0]]), kv, 2, {1: [0, 1]}, {1: 1})
    disp = e0 + 2 * t * np.cos(2 * np.pi * ks)
    assert np.allclose(out[0][0], disp, atol=1e-12) and np.allclose(out[0][1], 10.0, atol=1e-12)
    assert np.isclose(out[2][0], 10.0 - disp.max(), atol=1e-12)


def test_oracle_identity_overlap_and_edge_order():
    b = gd.Batch.from_data_list([gd.bulk_silicon(), gd.graphene(rep=(2, 2, 1), seed=1)])
    nao = 19
    N, E = b.num_nodes, b.edge_index.shape[1]
    g = torch.Generator().manual_seed(3)
    a = torch.randn(N, nao, nao, generator=g)
    hon = (0.5 * (a + a.transpose(1, 2))).reshape(N, -1)
    c = torch.randn(E, nao, nao, generator=g)
    hoff = (0.5 * (c + c[_global_inverse(b)].transpose(1, 2))).reshape(E, -1)
    son = torch.eye(nao).expand(N, nao, nao).reshape(N, -1)
    soff = torch.zeros(E, nao * nao)
    from hamgnn_b200.hamgnn_output import openmx_basis
    from hamgnn_b200.band import OPENMX_NUM_VALENCE
    _, _, basis = openmx_basis(nao)
    kv = np.random.default_rng(0).uniform(-1, 1, (2, 4, 3))
    args = dict(z=b.z.numpy(), batch=b.batch.numpy(), node_counts=b.node_counts.tolist(), k_vecs=kv, nao_max=nao, basis_def=basis,
                num_valence=OPENMX_NUM_VALENCE)
    be, wf, gap, hs, ksp = band_ref.band_energies(hon.numpy(), hoff.numpy(), son.numpy(), soff.numpy(), b.edge_index.numpy(),
                                                  nbr_shift=b.nbr_shift.numpy(), return_kspace=True, **args)
    # S = 1: bands = eigenvalues of the Hermitian H(k)
    row = 0
    for hk, sk in ksp:
        assert np.abs(hk - np.conj(np.swapaxes(hk, 1, 2))).max() < 1e-12
        ev = np.linalg.eigvalsh(hk)
        assert np.allclose(be[row:row + ev.shape[1]].T, ev, atol=1e-10)
        row += ev.shape[1]
    # a permutation of the edges inside each crystal changes nothing
    e_counts = np.bincount(b.batch.numpy()[b.edge_index[0].numpy()])
    perm = np.concatenate([off + np.random.default_rng(1).permutation(n) for off, n in zip(np.cumsum(e_counts) - e_counts, e_counts)])
    be2 = band_ref.band_energies(hon.numpy(), hoff.numpy()[perm], son.numpy(), soff.numpy()[perm], b.edge_index.numpy()[:, perm],
                                 nbr_shift=b.nbr_shift.numpy()[perm], **args)[0]
    assert np.allclose(be, be2, atol=1e-10)


def test_band_host_tables_cpu():
    """The host side of hgb_band_kspace (compact orbital index, edges grouped by atom pair) drives a CPU emulation of the two
    kernels to the oracle's H(k), S(k) -- on shuffled edges, so that the stable (i, j) grouping is exercised."""
    import hgb_kernel_emulator as EM
    from hamgnn_b200.band import BandEnergyHead, OPENMX_NUM_VALENCE
    from hamgnn_b200.hamgnn_output import openmx_basis
    nao = 19
    _, _, basis = openmx_basis(nao)
    g = gd.Batch.from_data_list([gd.mos2_monolayer(seed=2)])      # Mo (19 orbitals) and S (13): the mask matters
    N, E = g.num_nodes, g.edge_index.shape[1]
    gen = torch.Generator().manual_seed(9)
    perm = torch.randperm(E, generator=gen)
    src, dst, shift = g.edge_index[0][perm], g.edge_index[1][perm], g.nbr_shift[perm]
    hon, hoff = torch.randn(N, nao * nao, generator=gen), torch.randn(E, nao * nao, generator=gen)
    son, soff = torch.randn(N, nao * nao, generator=gen), torch.randn(E, nao * nao, generator=gen)
    kv = torch.rand(5, 3, generator=gen) - 0.5
    head = BandEnergyHead(nao, basis, OPENMX_NUM_VALENCE, 5)
    orb_index, n_orb, seg_ptr, order, n_segs = head.kspace_tables(src, dst, g.z)
    assert n_orb == sum(len(basis[int(z)]) for z in g.z) and n_segs == len(set(zip(src.tolist(), dst.tolist())))
    hk, sk = EM.emulate_band_kspace(hon, hoff, son, soff, nao, seg_ptr.numpy(), order.numpy(), src.numpy(), dst.numpy(),
                                    shift.numpy(), kv.numpy(), orb_index.numpy(), n_orb)
    # the oracle's dense scatter + orbital selection, without its eigen-solve (a random S is not positive definite)
    nk = kv.shape[0]
    dh = np.zeros((nk, N, N, nao, nao), dtype=np.complex128)
    ds = np.zeros_like(dh)
    for a in range(N):
        dh[:, a, a] += hon[a].double().numpy().reshape(nao, nao)
        ds[:, a, a] += son[a].double().numpy().reshape(nao, nao)
    for e in range(E):
        ph = np.exp(2j * np.pi * (kv.double().numpy() @ shift[e].double().numpy()))
        dh[:, int(src[e]), int(dst[e])] += ph[:, None, None] * hoff[e].double().numpy().reshape(nao, nao)
        ds[:, int(src[e]), int(dst[e])] += ph[:, None, None] * soff[e].double().numpy().reshape(nao, nao)
    keep = np.zeros((99, nao), dtype=bool)
    for z, idx in basis.items():
        keep[z, idx] = True
    keep = keep[g.z.numpy()].reshape(-1)
    dh = np.swapaxes(dh, -2, -3).reshape(nk, N * nao, N * nao)[:, keep][:, :, keep]
    ds = np.swapaxes(ds, -2, -3).reshape(nk, N * nao, N * nao)[:, keep][:, :, keep]
    assert np.abs(hk - dh).max() < 1e-10 and np.abs(sk - ds).max() < 1e-10


@pytest.mark.gpu
def test_band_head_matches_oracle_gpu():
    from hamgnn_b200.band import BandEnergyHead, OPENMX_NUM_VALENCE
    from hgb_testlib import SMALL_CFG, build_pair
    nao = 19
    pre, out, _o1, _o2 = build_pair(SMALL_CFG, nao_max=nao, add_H0=False)
    dev = torch.device("cuda:0")
    pre.to(dev); out.to(dev)
    b = gd.Batch.from_data_list([gd.bulk_silicon(), gd.graphene(rep=(2, 2, 1), seed=1)])
    son, soff = _pd_overlap(b, nao)
    kv = torch.tensor(np.random.default_rng(5).uniform(-0.5, 0.5, (2, 6, 3)), dtype=torch.float32)
    bd = gd.Batch(**b.to_dict()).to(dev)
    bd["Son"], bd["Soff"], bd["k_vecs"] = son.to(dev), soff.to(dev), kv.to(dev)
    with torch.no_grad():
        res = out(bd, pre(bd))
    H = res["hamiltonian"]
    on_row, off_row, _inv = out._row_maps(bd)
    hon, hoff = H[on_row].contiguous(), H[off_row].contiguous()
    for ctrl in (None, 3):
        head = BandEnergyHead(nao, out.basis_def, OPENMX_NUM_VALENCE, 6, ctrl)
        be, wf, gap, hs = head(hon, hoff, bd)
        torch.cuda.synchronize()
        ref = band_ref.band_energies(hon.cpu().numpy(), hoff.cpu().numpy(), son.numpy(), soff.numpy(), b.edge_index.numpy(),
                                     b.z.numpy(), b.batch.numpy(), b.node_counts.tolist(), b.nbr_shift.numpy(), kv.numpy(), nao,
                                     out.basis_def, OPENMX_NUM_VALENCE, ctrl, return_kspace=True)
        width = float(np.abs(ref[0]).max())
        assert be.shape == ref[0].shape
        assert float(np.abs(be.cpu().numpy() - ref[0]).max()) < 2e-4 * width, np.abs(be.cpu().numpy() - ref[0]).max()
        assert float(np.abs(gap.cpu().numpy() - ref[2]).max()) < 2e-4 * width
        assert wf.numel() == ref[1].size and hs.numel() == ref[3].size
        assert float(np.abs(hs.cpu().numpy() - ref[3]).max()) < 5e-4 * float(np.abs(ref[3]).max())
    # the reciprocal-space matrices themselves, crystal by crystal
    a0 = e0 = 0
    e_counts = torch.bincount(b.batch[b.edge_index[0]]).tolist()
    for c, (hk_ref, sk_ref) in enumerate(ref[4]):
        na, ne = int(b.node_counts[c]), int(e_counts[c])
        hk, sk = head.kspace(hon[a0:a0 + na], hoff[e0:e0 + ne], bd["Son"][a0:a0 + na], bd["Soff"][e0:e0 + ne],
                             (bd.edge_index[0][e0:e0 + ne] - a0).contiguous(), (bd.edge_index[1][e0:e0 + ne] - a0).contiguous(),
                             bd.nbr_shift[e0:e0 + ne], kv[c].to(dev), bd.z[a0:a0 + na])
        assert float(np.abs(hk.cpu().numpy() - hk_ref).max()) < 2e-6 * float(np.abs(hk_ref).max())
        assert float(np.abs(sk.cpu().numpy() - sk_ref).max()) < 2e-6 * float(np.abs(sk_ref).max())
        hk2, _ = head.kspace(hon[a0:a0 + na], hoff[e0:e0 + ne], bd["Son"][a0:a0 + na], bd["Soff"][e0:e0 + ne],
                             (bd.edge_index[0][e0:e0 + ne] - a0).contiguous(), (bd.edge_index[1][e0:e0 + ne] - a0).contiguous(),
                             bd.nbr_shift[e0:e0 + ne], kv[c].to(dev), bd.z[a0:a0 + na])
        assert torch.equal(torch.view_as_real(hk), torch.view_as_real(hk2)), "H(k) must be bit-reproducible"
        a0 += na; e0 += ne
    # through the module: calculate_band_energy=True fills the result (and data.band_energy from data.Hon / Hoff when present)
    from hamgnn_b200.hamgnn_output import HamGNNPlusPlusOut
    out2 = HamGNNPlusPlusOut(pre.irreps_node_features, pre.irreps_node_features, nao_max=nao, soc_switch=False, ham_only=True,
                             add_H0=False, calculate_band_energy=True, num_k=6, band_num_control=3).to(dev)
    out2.load_state_dict(out.state_dict())
    with torch.no_grad():
        res2 = out2(bd, pre(bd))
    assert torch.equal(res2["hamiltonian"], H)
    assert float((res2["band_energy"] - be).abs().max()) < 1e-4 * width and res2["band_gap"].shape == (2,)
