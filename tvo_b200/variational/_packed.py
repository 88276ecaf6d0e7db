"""Bit-packed state storage with the reference's uint8 (N, S, H) face.

The reference stores K as uint8 (N,S,H) (TVOVariationalStates.py:41-43): 8x the bytes the
information needs, and 131 GB at BASELINE config 5.  Here the canonical store is an int64 tensor
(N, S, W64) holding uint64 bit patterns (bit h%64 of word h//64 is s_h); `PackedStates` presents it
with the reference's shape / indexing semantics by (un)packing on the device on demand.
"""
import torch as to

from tvo_b200 import _lib


def n_words(H: int) -> int:
    return (int(H) + 63) // 64


def pack(K_u8: to.Tensor) -> to.Tensor:
    """uint8 (..., H) on the device -> int64 (..., W64) bit patterns (kernel pack_states)."""
    assert K_u8.dtype == to.uint8, "states must be uint8"
    H = K_u8.shape[-1]
    K_u8 = K_u8.contiguous()
    n = K_u8.numel() // H if H else 0
    out = to.empty(K_u8.shape[:-1] + (n_words(H),), dtype=to.int64, device=K_u8.device)
    _lib.call("tvo_pack_states", _lib.ptr(K_u8), _lib.ptr(out), n, H, _lib.stream())
    return out


def unpack(packed: to.Tensor, H: int) -> to.Tensor:
    """int64 (..., W64) -> uint8 (..., H) (kernel unpack_states)."""
    packed = packed.contiguous()
    n = packed.numel() // n_words(H)
    out = to.empty(packed.shape[:-1] + (H,), dtype=to.uint8, device=packed.device)
    _lib.call("tvo_unpack_states", _lib.ptr(packed), _lib.ptr(out), n, H, _lib.stream())
    return out


class PackedStates:
    """uint8 (N, S, H) view of a packed int64 (N or 1, S, W64) tensor.

    `broadcast=True` means one state set shared by all N datapoints (FullEM's stride-0 expand,
    fullem.py:48).  Indexing along dim 0 returns *unpacked* uint8 tensors like the reference's K;
    `rows(idx)` returns the packed rows for kernels."""

    def __init__(self, packed: to.Tensor, H: int, N: int = None):
        assert packed.dtype == to.int64 and packed.dim() == 3
        self.packed, self.H = packed, int(H)
        self.N = packed.shape[0] if N is None else int(N)
        self.broadcast = packed.shape[0] == 1 and self.N != 1
        assert self.broadcast or packed.shape[0] == self.N

    # --- tensor-like face -------------------------------------------------------------------
    @property
    def shape(self):
        return to.Size((self.N, self.packed.shape[1], self.H))

    @property
    def device(self):
        return self.packed.device

    dtype = to.uint8
    ndim = 3

    def dim(self):
        return 3

    def size(self, d=None):
        return self.shape if d is None else self.shape[d]

    def numel(self):
        return self.N * self.packed.shape[1] * self.H

    def rows(self, idx=None) -> to.Tensor:
        """Packed rows (B, S, W64) for datapoints idx (all if None)."""
        if self.broadcast:
            return self.packed.expand(self.N if idx is None else self._count(idx), -1, -1)
        return self.packed if idx is None else self.packed[idx]

    def unpack(self) -> to.Tensor:
        u = unpack(self.packed, self.H)
        return u.expand(self.N, -1, -1) if self.broadcast else u

    def _count(self, first):
        if isinstance(first, to.Tensor):
            return int(first.sum()) if first.dtype == to.bool else first.numel()
        if isinstance(first, slice):
            return len(range(*first.indices(self.N)))
        return len(first)

    def __getitem__(self, key):
        first, rest = (key[0], tuple(key[1:])) if isinstance(key, tuple) else (key, ())
        if isinstance(first, int):
            u = unpack(self.packed[0 if self.broadcast else first], self.H)  # (S, H)
            return u[rest] if rest else u
        if self.broadcast:
            u = unpack(self.packed, self.H).expand(self._count(first), -1, -1)
        else:
            if isinstance(first, to.Tensor):
                first = first.to(self.packed.device)
            u = unpack(self.packed[first], self.H)
        return u[(slice(None),) + rest] if rest else u

    def __setitem__(self, key, value):
        assert not self.broadcast, "cannot assign into a broadcast state set"
        if isinstance(key, tuple):
            u = self.unpack()
            u[key] = value
            self.packed.copy_(pack(u))
            return
        if isinstance(value, PackedStates):
            self.packed[key] = value.packed
        else:
            value = value.to(device=self.packed.device, dtype=to.uint8)
            tgt = self.packed[key]
            self.packed[key] = pack(value.expand(tgt.shape[:-1] + (self.H,)).contiguous())

    def clone(self):
        return self.unpack().clone()

    def to(self, *a, **kw):
        return self.unpack().to(*a, **kw)

    def cpu(self):
        return self.unpack().cpu()

    def numpy(self):
        return self.unpack().cpu().numpy()

    def sum(self, *a, **kw):
        return self.unpack().sum(*a, **kw)

    def any(self, *a, **kw):
        return self.unpack().any(*a, **kw)

    def all(self, *a, **kw):
        return self.unpack().all(*a, **kw)

    def __eq__(self, other):
        return self.unpack() == (other.unpack() if isinstance(other, PackedStates) else other)

    def __ne__(self, other):
        return self.unpack() != (other.unpack() if isinstance(other, PackedStates) else other)

    __hash__ = None

    def __len__(self):
        return self.N

    def __repr__(self):
        return f"PackedStates(shape={tuple(self.shape)}, device={self.device}, broadcast={self.broadcast})"


def as_packed(states, H=None):
    """(packed int64 (n,S,W64), H, k_stride_n_is_zero) from uint8 tensor / PackedStates."""
    if isinstance(states, PackedStates):
        return states.packed, states.H, states.broadcast
    assert isinstance(states, to.Tensor)
    if states.dtype == to.int64 and H is not None:
        return states, H, False
    assert states.dtype == to.uint8 and states.dim() == 3, "states must be uint8 (N,S,H)"
    H = states.shape[2]
    if states.shape[0] > 1 and states.stride(0) == 0:  # FullEM-style expanded view
        return pack(states[:1]), H, True
    return pack(states), H, False
