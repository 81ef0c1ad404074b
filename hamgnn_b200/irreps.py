"""Irreps bookkeeping for the B200 engine (host side).

String-compatible with the `e3nn.o3.Irreps` objects the reference passes around
(/root/reference/hamgnn/models/hamgnn_conv.py:144, hamgnn/main.py:235-236): `str(Irreps(...))`
round-trips, `Irreps(str(e3nn_irreps))` works, and `sort()` / `simplify()` follow e3nn's rules
(sort key = (l, p, position) with p=-1 before p=+1; simplify merges adjacent equal irreps only) because
those rules fix the flat weight layouts of the reference's TensorProduct / Linear parameters
(hamgnn/nn/message_passing.py:156-167).
"""
from __future__ import annotations

from typing import Iterable, List, NamedTuple, Tuple


class Ir(NamedTuple):
    l: int
    p: int  # +1 even, -1 odd

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def __str__(self):
        return f"{self.l}{'e' if self.p > 0 else 'o'}"

    __repr__ = __str__

    @staticmethod
    def parse(s) -> "Ir":
        if isinstance(s, Ir):
            return s
        if isinstance(s, (tuple, list)):
            l, p = s
            return Ir(int(l), int(p))
        s = str(s).strip()
        return Ir(int(s[:-1]), 1 if s[-1] == "e" else -1)

    def product(self, other: "Ir") -> List["Ir"]:
        p = self.p * other.p
        return [Ir(l, p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]


class MulIr(NamedTuple):
    mul: int
    ir: Ir

    @property
    def dim(self) -> int:
        return self.mul * self.ir.dim


class Irreps:
    __slots__ = ("items",)

    def __init__(self, spec=None):
        items: List[MulIr] = []
        if spec is None:
            pass
        elif isinstance(spec, Irreps):
            items = list(spec.items)
        elif isinstance(spec, str) or hasattr(spec, "__str__") and not isinstance(spec, (list, tuple)):
            for tok in str(spec).split("+"):
                tok = tok.strip()
                if not tok:
                    continue
                if "x" in tok:
                    m, ir = tok.split("x")
                    items.append(MulIr(int(m), Ir.parse(ir)))
                else:
                    items.append(MulIr(1, Ir.parse(tok)))
        else:
            for it in spec:
                if isinstance(it, MulIr):
                    items.append(it)
                elif isinstance(it, Ir):
                    items.append(MulIr(1, it))
                else:
                    m, ir = it
                    items.append(MulIr(int(m), Ir.parse(ir)))
        self.items = items

    # -- container protocol
    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]

    def __eq__(self, other):
        return isinstance(other, Irreps) and self.items == other.items

    def __str__(self):
        return "+".join(f"{m}x{ir}" for m, ir in self.items)

    __repr__ = __str__

    def __add__(self, other):
        return Irreps(self.items + Irreps(other).items)

    def __mul__(self, n: int):
        return Irreps(self.items * int(n))

    __rmul__ = __mul__

    # -- sizes
    @property
    def dim(self) -> int:
        return sum(x.dim for x in self.items)

    @property
    def num_irreps(self) -> int:
        return sum(x.mul for x in self.items)

    @property
    def lmax(self) -> int:
        return max(x.ir.l for x in self.items)

    def offsets(self) -> List[int]:
        """Start column of every slot in the flattened feature row."""
        out, o = [], 0
        for x in self.items:
            out.append(o)
            o += x.dim
        return out

    def channel_offsets(self) -> List[int]:
        out, o = [], 0
        for x in self.items:
            out.append(o)
            o += x.mul
        return out

    # -- e3nn-compatible transforms
    def sort(self) -> Tuple["Irreps", Tuple[int, ...], Tuple[int, ...]]:
        order = sorted(range(len(self.items)), key=lambda i: (self.items[i].ir.l, self.items[i].ir.p, i))
        perm = [0] * len(order)
        for new, old in enumerate(order):
            perm[old] = new
        return Irreps([self.items[i] for i in order]), tuple(perm), tuple(order)

    def simplify(self) -> "Irreps":
        out: List[MulIr] = []
        for m, ir in self.items:
            if out and out[-1].ir == ir:
                out[-1] = MulIr(out[-1].mul + m, ir)
            elif m > 0:
                out.append(MulIr(m, ir))
        return Irreps(out)

    def scaled(self, factor) -> "Irreps":
        """hamgnn/utils/irreps_utils.py:67-79 (scale_irreps)."""
        return Irreps([MulIr(max(1, int(m * factor)), ir) for m, ir in self.items])
