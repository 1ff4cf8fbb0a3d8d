"""Python mirror of the plugin-side host logic (include/egpu_plugin.h), used by the tests
the way a Go handler would use the cgo binding: device-ID codec and
GetPreferredAllocation for one container request (pkg/plugins/base.go:94-96).
All work happens in the C/CUDA library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

RESOURCE_CORE = 0  # elasticgpu.io/gpu-core
RESOURCE_MEM = 1   # elasticgpu.io/gpu-memory


def format_device_id(gpu: int, unit: int) -> str:
    buf = C.create_string_buffer(40)
    n = L.load().egpu_device_id_format(int(gpu), int(unit), buf, 40)
    if n < 0:
        raise L.EgpuError(n, "egpu_device_id_format")
    return buf.value.decode()


def parse_device_id(s: str) -> tuple[int, int]:
    g, u = C.c_int32(), C.c_int64()
    rc = L.load().egpu_device_id_parse(s.encode(), C.byref(g), C.byref(u))
    if rc != L.OK:
        raise L.EgpuError(rc, f"egpu_device_id_parse({s!r})")
    return g.value, u.value


def _strs(ids):
    arr = (C.c_char_p * max(1, len(ids)))()
    keep = [s.encode() for s in ids]
    for i, b in enumerate(keep):
        arr[i] = b
    return arr, keep


def preferred_allocation(alloc, available: list[str], must_include: list[str], allocation_size: int,
                         resource: int = RESOURCE_CORE) -> tuple[list[str], int]:
    """ContainerPreferredAllocationRequest -> (deviceIDs, gpu index)."""
    av, _k1 = _strs(available)
    mu, _k2 = _strs(must_include)
    out = np.full(max(1, allocation_size), -1, dtype=np.int32)
    gpu = C.c_int32(-1)
    rc = L.load().egpu_preferred_allocation(alloc.handle, av, len(available), mu, len(must_include), int(allocation_size),
                                            int(resource), C.c_void_p(out.ctypes.data), C.byref(gpu))
    if rc != L.OK:
        raise L.EgpuError(rc, "egpu_preferred_allocation")
    return [available[p] for p in out[:allocation_size]], gpu.value
