"""Host side of egpu_table_restore on CPU: the JSON reader, the symlink parsers and the
flattening of elastic-gpu-agent_b200/csrc/egpu_restore.cc, linked against a recording stand-in
for egpu_table_restore_flat (tests/restore_host_harness.cc).  What the device is handed must be
what the oracle's own reading of the same records gives."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import restore_py as R

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "restore_host_harness.so")  # test artefact, not a product library
G = json.load(open(os.path.join(HERE, "golden", "restore_records.json")))


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O3", "-Wall", "-Werror", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "elastic-gpu-agent_b200", "csrc", "egpu_restore.cc"), os.path.join(HERE, "restore_host_harness.cc"),
           "-o", SO]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(SO)
    lib.cap_n.restype = C.c_int64
    lib.cap_n.argtypes = [C.c_int]
    lib.cap_ptr.restype = C.c_void_p
    lib.cap_ptr.argtypes = [C.c_int]
    lib.cap_error.restype = C.c_char_p
    lib.egpu_table_restore.restype = C.c_int
    return lib


def call(lib, records, links, D=4, flags=3):
    n, nl = len(records), len(links)
    keys = (C.c_char_p * max(1, n))(*[k for k, _ in records])
    vals = (C.c_char_p * max(1, n))(*[v for _, v in records])
    klen = (C.c_int64 * max(1, n))(*[len(k) for k, _ in records])
    vlen = (C.c_int64 * max(1, n))(*[len(v) for _, v in records])
    names = (C.c_char_p * max(1, nl))(*[a.encode() for a, _ in links])
    targets = (C.c_char_p * max(1, nl))(*[b.encode() for _, b in links])
    capc = (C.c_int32 * D)(*([100] * D))
    capm = (C.c_int32 * D)(*([1000] * D))
    table = (C.c_int32 * (3 * D))()
    counts = (C.c_int64 * 5)()
    rstat = (C.c_int32 * max(1, n))()
    rc = lib.egpu_table_restore(C.c_void_p(1), keys, klen, vals, vlen, C.c_int64(n), names, targets, C.c_int64(nl), capc, capm,
                                C.c_int32(D), C.c_int(flags), table, counts, rstat)
    return rc, list(counts), list(rstat)[:n]


def captured(lib):
    def arr(what, dtype):
        n = lib.cap_n(what)
        if n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(lib.cap_ptr(what), C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()
    flat = C.string_at(lib.cap_ptr(0), lib.cap_n(0)) if lib.cap_n(0) else b""
    hash8 = C.string_at(lib.cap_ptr(3), lib.cap_n(3)) if lib.cap_n(3) else b""
    return dict(flat=flat, id_off=arr(1, np.int64), set_off=arr(2, np.int64), hash8=hash8, resource=arr(4, np.int32),
                link_off=arr(5, np.int64), link_gpu=arr(6, np.int32))


def expected_entries(records, links):
    """The oracle's reading: one entry per container, in record order then first-appearance key order."""
    link = R.parse_links(links)
    out = []
    for key, val in records:
        obj = json.loads(val.decode()) or {}
        for _c, dev in obj.items():
            is_null = dev is None
            dev = dev or {}
            res = R._field(dev, "ResourceName") or ""
            h = R._field(dev, "Hash") or ""
            ids = R._field(dev, "List") or []
            code = 0 if (res == R.CORE or is_null) else 1 if res == R.MEM else -1
            if code == -1 or is_null:
                ids = []
            lk = link.get(h, {})
            gpus = [lk.get(i, -1) for i in range(max(lk) + 1)] if lk else []
            out.append((ids, (h + "????????")[:8] if len(h) == 8 else "????????", code, gpus))
    return out


def test_golden_records_are_flattened_as_the_oracle_reads_them(harness):
    records = [(k.encode(), v.encode()) for k, v in G["records"]]
    links = [tuple(l) for l in G["links"]]
    rc, counts, rstat = call(harness, records, links)
    assert rc == 0 and harness.cap_n(8) == 3
    cap = captured(harness)
    exp = expected_entries(records, links)
    assert len(cap["resource"]) == len(exp) == 12
    for q, (ids, h8, code, gpus) in enumerate(exp):
        lo, hi = cap["set_off"][q], cap["set_off"][q + 1]
        got_ids = [cap["flat"][cap["id_off"][i]:cap["id_off"][i + 1]].decode() for i in range(lo, hi)]
        assert got_ids == ids, q
        assert cap["hash8"][8 * q:8 * q + 8].decode() == h8, q
        assert cap["resource"][q] == code, q
        assert cap["link_gpu"][cap["link_off"][q]:cap["link_off"][q + 1]].tolist() == gpus, q
    # the per-status counts and the per-record worst status are derived from what the device reports
    # (the stand-in reports q % 5): counts of 12 entries, worst per record
    assert counts == [3, 3, 2, 2, 2]
    per_entry = [q % 5 for q in range(12)]
    owners = [r for r, (_, v) in enumerate(records) for _ in (json.loads(v.decode()) or {})]
    assert rstat == [max(s for s, o in zip(per_entry, owners) if o == r) for r in range(len(records))]


@pytest.mark.parametrize("key,val", [tuple(x) for x in G["parse_errors"]])
def test_parse_errors_abort_before_the_device_is_touched(harness, key, val):
    before = harness.cap_n(7)
    good = tuple(x.encode() for x in G["records"][0])
    rc, _, _ = call(harness, [good, (key.encode(), val.encode())], [])
    assert rc == -7 and harness.cap_n(7) == before
    assert b"record 1" in harness.cap_error()


@pytest.mark.parametrize("val,ids", [
    (b'{"c":{"Hash":"00000000","List":["1-05","2-07"],"ResourceName":"elasticgpu.io/gpu-memory"}}', ["1-05", "2-07"]),
    (b' \n{"c" : {"hash":"00000000" , "LIST":[ "1-05" ,\t"2-07" ] , "resourceName" : "elasticgpu.io/gpu-memory", "x":{"y":[1,2,{"z":null}]}} } ',
     ["1-05", "2-07"]),
    (b'{"c":{"Hash":"00000000","List":["1-\\u0030\\u0035","2-07"],"ResourceName":"elasticgpu.io\\/gpu-memory"}}', ["1-05", "2-07"]),
    (b'{"c":{"List":["9-99"],"ResourceName":"elasticgpu.io/gpu-memory"},"c":{"Hash":"00000000","List":["1-05","2-07"],'
     b'"ResourceName":"elasticgpu.io/gpu-memory"}}', ["1-05", "2-07"]),           # duplicate key: the last one wins
    (b'null', None),
    (b'{}', None),
])
def test_json_leniency_matches_encoding_json(harness, val, ids):
    rc, _, _ = call(harness, [(b"ns/pod", val)], [])
    assert rc == 0
    cap = captured(harness)
    if ids is None:
        assert len(cap["resource"]) == 0
        return
    assert len(cap["resource"]) == 1 and cap["resource"][0] == 1
    got = [cap["flat"][cap["id_off"][i]:cap["id_off"][i + 1]].decode() for i in range(len(cap["id_off"]) - 1)]
    assert got == ids


@pytest.mark.parametrize("name,target,ok", [
    ("elastic-gpu-00000000-0", "/dev/nvidia3", True), ("00000000-0", "/dev/nvidia3", True),
    ("elastic-gpuctl-00000000-0", "/dev/nvidia3", False), ("elastic-gpu-00000000-0", "/dev/nvidiactl", False),
    ("elastic-gpu-00000000-0", "/dev/nvidia", False), ("elastic-gpu-00000000-x", "/dev/nvidia3", False),
    ("elastic-gpu-00000000", "/dev/nvidia3", False), ("elastic-gpu-00000000-0", "/dev/nvidia3x", False),
    ("elastic-gpu-00000000-00000000000", "/dev/nvidia3", False), ("nvidia0", "/dev/nvidia0", False),
])
def test_symlink_parsers(harness, name, target, ok):
    val = b'{"c":{"Hash":"00000000","List":["1-05"],"ResourceName":"elasticgpu.io/gpu-core"}}'
    rc, _, _ = call(harness, [(b"ns/pod", val)], [(name, target)])
    assert rc == 0
    cap = captured(harness)
    assert cap["link_gpu"].tolist() == ([3] if ok else [])
    assert bool(R.parse_links([(name, target)]).get("00000000")) == ok
