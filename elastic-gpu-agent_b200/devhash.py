"""Python mirror of include/egpu_devhash.h: the reference's device-set identity
(types.NewDevice / hash / Equals, pkg/types/device.go:17-54) and the search of
KubeletDeviceLocator.Locate (pkg/kube/locator.go:62-90), batched on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _flatten(sets: list[list[str]]):
    flat = bytearray()
    id_off = [0]
    set_off = [0]
    for ids in sets:
        for s in ids:
            flat += s.encode()
            id_off.append(len(flat))
        set_off.append(len(id_off) - 1)
    return (bytes(flat), np.asarray(id_off, dtype=np.int64), np.asarray(set_off, dtype=np.int64))


def flatten(sets: list[list[str]]):
    """(ids_flat bytes, id_offsets int64[n_ids+1], set_offsets int64[n_sets+1]) — the C ABI's input layout."""
    return _flatten(sets)


def device_hashes_flat(alloc, flat: bytes, id_off: np.ndarray, set_off: np.ndarray):
    """egpu_device_hash_batch on already flattened input; returns the 8-hex-digit hashes."""
    n_sets = len(set_off) - 1
    out = C.create_string_buffer(9 * n_sets)
    rc = L.load().egpu_device_hash_batch(alloc.handle, C.c_char_p(flat), C.c_void_p(id_off.ctypes.data), len(id_off) - 1,
                                         C.c_void_p(set_off.ctypes.data), n_sets, out, None)
    if rc != L.OK:
        raise L.EgpuError(rc, "egpu_device_hash_batch")
    return [out.raw[9 * i:9 * i + 8].decode() for i in range(n_sets)]


def locate_flat(alloc, flat: bytes, id_off: np.ndarray, set_off: np.ndarray) -> int:
    """egpu_device_locate on flattened input (set 0 = request); candidate index or -1."""
    m = C.c_int64(-1)
    rc = L.load().egpu_device_locate(alloc.handle, C.c_char_p(flat), C.c_void_p(id_off.ctypes.data), len(id_off) - 1,
                                     C.c_void_p(set_off.ctypes.data), len(set_off) - 1, C.byref(m))
    if rc != L.OK:
        raise L.EgpuError(rc, "egpu_device_locate")
    return int(m.value) - 1 if m.value >= 1 else -1


def device_hashes(alloc, sets: list[list[str]], want_digest: bool = False):
    """Device.Hash (8 hex digits) of every ID list; optionally the full SHA-256 digests too."""
    flat, id_off, set_off = _flatten(sets)
    n_sets = len(sets)
    out = C.create_string_buffer(9 * n_sets)
    dg = np.zeros(32 * n_sets, dtype=np.uint8)
    rc = L.load().egpu_device_hash_batch(alloc.handle, C.c_char_p(flat), C.c_void_p(id_off.ctypes.data), len(id_off) - 1,
                                         C.c_void_p(set_off.ctypes.data), n_sets, out,
                                         C.c_void_p(dg.ctypes.data) if want_digest else None)
    if rc != L.OK:
        raise L.EgpuError(rc, "egpu_device_hash_batch")
    hashes = [out.raw[9 * i:9 * i + 8].decode() for i in range(n_sets)]
    if want_digest:
        return hashes, [bytes(dg[32 * i:32 * i + 32]) for i in range(n_sets)]
    return hashes


def device_hash(alloc, ids: list[str]) -> str:
    arr = (C.c_char_p * max(1, len(ids)))(*[s.encode() for s in ids])
    out = C.create_string_buffer(9)
    rc = L.load().egpu_device_hash(alloc.handle, arr, len(ids), out)
    if rc != L.OK:
        raise L.EgpuError(rc, "egpu_device_hash")
    return out.value.decode()


def locate(alloc, request: list[str], candidates: list[list[str]]) -> int:
    """Index into `candidates` of the first list equal (as a sorted multiset) to `request`, or -1."""
    flat, id_off, set_off = _flatten([request] + candidates)
    m = C.c_int64(-1)
    rc = L.load().egpu_device_locate(alloc.handle, C.c_char_p(flat), C.c_void_p(id_off.ctypes.data), len(id_off) - 1,
                                     C.c_void_p(set_off.ctypes.data), len(candidates) + 1, C.byref(m))
    if rc != L.OK:
        raise L.EgpuError(rc, "egpu_device_locate")
    return int(m.value) - 1 if m.value >= 1 else -1
