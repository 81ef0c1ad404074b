"""`graph_data_module` / `NPZGraphDataset`: the data side of the drop-in boundary.

Mirrors the public surface of /root/reference/hamgnn/data/graph_data.py that `hamgnn/main.py:133-160` touches --
constructor keywords, `prepare_data()`, `setup(stage)`, `{train,val,test}_dataloader()`, `save_split()`, the
`train_data/val_data/test_data` attributes, the seed-42 shuffled split rule (:367-380) and the optional
`split_file` with `train_idx/val_idx/test_idx` (:360-366) -- for in-memory graph lists and `graph_data.npz`
files.  Batches are collated with PyG's rule (`hamgnn_b200.graph_data.Batch.from_data_list`).  LMDB input
(`LMDBGraphDataset`, :25-94) belongs to the storage side that SURVEY.md section 8 leaves out of scope: it raises.

When `pytorch_lightning` is importable the module derives from `LightningDataModule`, so a Lightning `Trainer`
accepts it; otherwise it is a plain object with the same methods.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, Subset

from .graph_data import Batch, Data, _install_pyg_stub

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as _pl
    _Base = _pl.LightningDataModule
except Exception:  # noqa: BLE001
    _Base = object


def _revive(g) -> Data:
    """One stored graph -> `Data` (pickled PyG Data, our Data, or a dict of numpy arrays; reference :149-167)."""
    if isinstance(g, Data):
        return g
    if isinstance(g, dict):
        if "edge_index" not in g:
            raise ValueError("graph dict without 'edge_index'")
        return Data(**{k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in g.items()})
    if hasattr(g, "keys"):   # a real torch_geometric Data
        keys = g.keys() if callable(g.keys) else g.keys
        return Data(**{k: g[k] for k in keys})
    raise TypeError(f"cannot interpret stored graph of type {type(g)}")


class NPZGraphDataset(Dataset):
    """`np.savez(path, graph={idx: Data})` reader with an index subset (reference :96-185)."""

    def __init__(self, npz_path: str, indices: Optional[List[int]] = None, transform: Optional[Callable] = None,
                 preload: int = 0):
        super().__init__()
        self.npz_path, self.transform = npz_path, transform
        _install_pyg_stub()
        try:
            with np.load(npz_path, allow_pickle=True) as f:
                if "graph" in f:
                    payload = f["graph"].item()
                    self.data_list = list(payload.values()) if isinstance(payload, dict) else list(payload)
                else:
                    self.data_list = [f[k] for k in f.keys()]
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Failed to load NPZ file: {e}") from e
        self.total_length = len(self.data_list)
        self.indices = list(indices) if indices is not None else list(range(self.total_length))
        self._cache = {}
        for real in self.indices[:max(0, int(preload))]:
            self._cache[real] = _revive(self.data_list[real])

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, idx):
        if isinstance(idx, list):
            return [self[i] for i in idx]
        real = self.indices[idx]
        g = self._cache.get(real)
        if g is None:
            g = _revive(self.data_list[real])
        return self.transform(g) if self.transform is not None else g


def _collate(graphs: Sequence[Data]) -> Batch:
    return Batch.from_data_list(list(graphs))


class graph_data_module(_Base):
    def __init__(self, dataset: Union[list, tuple, np.ndarray, str] = None, train_ratio: float = 0.6,
                 val_ratio: float = 0.2, test_ratio: float = 0.2, batch_size: int = 64, val_batch_size: int = None,
                 test_batch_size: int = None, split_file: str = None, num_workers: int = 4, prefetch_factor: int = 2,
                 cache_size: int = 100, transform: Callable = None, persistent_workers: bool = True, preload: int = 0,
                 test_mode: bool = False, data_format: str = "auto"):
        super().__init__()
        self.dataset_input = dataset
        self.train_ratio, self.val_ratio, self.test_ratio = train_ratio, val_ratio, test_ratio
        self.batch_size = batch_size
        self.val_batch_size = val_batch_size or batch_size
        self.test_batch_size = test_batch_size or self.val_batch_size
        self.split_file = split_file
        self.num_workers, self.prefetch_factor, self.cache_size = num_workers, prefetch_factor, cache_size
        self.transform, self.persistent_workers, self.preload = transform, persistent_workers, preload
        self.test_mode, self.data_format = test_mode, data_format

    # ------------------------------------------------------------------------------------------------
    def _resolve_format(self):
        if not isinstance(self.dataset_input, str):
            return "list"
        if self.data_format == "auto":
            if self.dataset_input.endswith(".npz"):
                self.data_format = "npz"
            elif self.dataset_input.endswith(".lmdb"):
                self.data_format = "lmdb"
            else:
                raise ValueError(f"The data format cannot be determined from the file extension: {self.dataset_input}")
        if self.data_format == "lmdb":
            raise NotImplementedError("LMDB graph stores are outside the B200 hot path (SURVEY.md section 8): convert "
                                      "with the reference's tools/npz_to_lmdb.py counterpart or pass graph_data.npz")
        if self.data_format != "npz":
            raise ValueError(f"Unsupported data format: {self.data_format}")
        return "npz"

    def prepare_data(self):
        if self._resolve_format() == "npz":
            if not os.path.exists(self.dataset_input):
                raise FileNotFoundError(f"The data path does not exist: {self.dataset_input}")
            print(f"Found {self.get_dataset_length()} graphs in the NPZ file")

    def get_dataset_length(self) -> int:
        if self._resolve_format() == "npz":
            with np.load(self.dataset_input, allow_pickle=True) as f:
                return len(f["graph"].item()) if "graph" in f else len(f.keys())
        return len(self.dataset_input)

    def _subset(self, indices):
        if self._resolve_format() == "npz":
            return NPZGraphDataset(self.dataset_input, indices=indices, transform=self.transform, preload=self.preload)
        return Subset(self.dataset_input, indices=list(indices))

    def setup(self, stage: Optional[str] = None):
        n = self.get_dataset_length()
        if self.test_mode:
            self.train_data, self.val_data, self.test_data = self._subset([]), self._subset([]), self._subset(range(n))
            return
        if self.split_file is not None and os.path.exists(self.split_file):
            sp = np.load(self.split_file)
            tr, va, te = (sp[k].tolist() for k in ("train_idx", "val_idx", "test_idx"))
        else:
            idx = list(range(n))
            np.random.RandomState(seed=42).shuffle(idx)          # reference split rule (:367-380)
            n_tr, n_va = round(self.train_ratio * n), round(self.val_ratio * n)
            tr, va, te = idx[:n_tr], idx[n_tr:n_tr + n_va], idx[n_tr + n_va:]
            if self.split_file is not None:
                np.savez(self.split_file, train_idx=np.array(tr), val_idx=np.array(va), test_idx=np.array(te))
        if stage in ("fit", None):
            self.train_data, self.val_data = self._subset(tr), self._subset(va)
        if stage in ("test", None):
            self.test_data = self._subset(te)

    def _loader(self, ds, batch_size, shuffle):
        kw = dict(batch_size=batch_size, shuffle=shuffle, pin_memory=torch.cuda.is_available(), collate_fn=_collate,
                  num_workers=self.num_workers)
        if self.num_workers > 0:
            kw.update(prefetch_factor=self.prefetch_factor, persistent_workers=self.persistent_workers)
        return DataLoader(ds, **kw)

    def train_dataloader(self):
        return self._loader(self.train_data, self.batch_size, True)

    def val_dataloader(self):
        return self._loader(self.val_data, self.val_batch_size, False)

    def test_dataloader(self):
        return self._loader(self.test_data, self.test_batch_size, False)

    def save_split(self, split_file: str):
        if not all(hasattr(self, a) for a in ("train_data", "val_data", "test_data")):
            raise RuntimeError("Dataset has not been set up yet. Call setup() first.")
        np.savez(split_file, **{f"{n}_idx": np.array(getattr(self, f"{n}_data").indices) for n in ("train", "val", "test")})
