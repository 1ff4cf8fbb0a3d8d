"""GetPreferredAllocation's host logic (elastic-gpu-agent_b200/csrc/egpu_plugin.cc) on CPU: the
file is linked against a recording stand-in for egpu_bestfit_query
(tests/plugin_host_harness.cc).  Checked here: the availability table built from the ID strings,
must-include pinning, the request handed to the scan, and the IDs chosen for the scan's answer -
the answer itself is scripted from the oracle (the CUDA scan is checked in the gpu tests)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "plugin_host_harness.so")  # test artefact, not a product library
CORE, MEM = 0, 1
CORE_MAX, MEM_MAX = 100, (1 << 18) - 1


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "elastic-gpu-agent_b200", "csrc", "egpu_plugin.cc"), os.path.join(HERE, "plugin_host_harness.cc"),
           "-o", SO]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    h = C.CDLL(SO)
    h.egpu_preferred_allocation.restype = C.c_int
    return h


def ids(gpu, units):
    return ["%d-%02d" % (gpu, u) for u in units]


def ask(lib, oracle_c, available, must, size, resource):
    """Runs the host logic twice: once to learn the table/request it builds (scripted answer -1),
    then with the oracle's answer for that table.  Returns (rc, chosen ids, gpu, table, request)."""
    def run(answer):
        lib.stub_script_answer(C.c_int32(answer))
        av = (C.c_char_p * max(1, len(available)))(*[s.encode() for s in available])
        mu = (C.c_char_p * max(1, len(must)))(*[s.encode() for s in must])
        pos = (C.c_int32 * max(1, size))()
        gpu = C.c_int32(-1)
        rc = lib.egpu_preferred_allocation(C.c_void_p(1), av, C.c_int64(len(available)), mu, C.c_int64(len(must)), C.c_int32(size),
                                           C.c_int(resource), pos, C.byref(gpu))
        return rc, [available[p] for p in list(pos)[:size]] if rc == 0 else [], gpu.value
    before = (C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32())
    lib.stub_request(*[C.byref(x) for x in before])
    rc, chosen, gpu = run(-1)
    after = (C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32())
    lib.stub_request(*[C.byref(x) for x in after])
    if after[3].value == before[3].value:       # rejected before the device was asked
        return rc, [], -1, None, None
    fc, fm = (C.c_int32 * 64)(), (C.c_int32 * 64)()
    D = lib.stub_table(fc, fm)
    table = (list(fc)[:D], list(fm)[:D])
    req = (after[0].value, after[1].value)
    tfc, tfm = np.asarray(table[0], np.int32), np.asarray(table[1], np.int32)   # (kept alive across the call)
    answer = oracle_c.load().oracle_pick_one(C.c_void_p(tfc.ctypes.data), C.c_void_p(tfm.ctypes.data), D, req[0], req[1])
    rc, chosen, gpu = run(answer)
    return rc, chosen, gpu, table, req


def test_table_request_and_choice(lib, oracle_c):
    # GPU 0: 60 core units free, GPU 1: 30, GPU 2: 25 (units shuffled on purpose)
    available = ids(0, range(40, 100)) + ids(2, range(75, 100))[::-1] + ids(1, range(0, 30))
    rc, chosen, gpu, table, req = ask(lib, oracle_c, available, [], 25, CORE)
    assert rc == 0 and table == ([60, 30, 25], [MEM_MAX] * 3) and req == (25, 1)
    assert gpu == 2 and chosen == ids(2, range(75, 100))                       # exact fit, ascending units
    rc, chosen, gpu, table, req = ask(lib, oracle_c, available, [], 26, CORE)
    assert rc == 0 and gpu == 1 and chosen == ids(1, range(0, 26))
    rc, chosen, gpu, table, req = ask(lib, oracle_c, available, [], 31, CORE)
    assert rc == 0 and gpu == 0 and chosen == ids(0, range(40, 71))
    rc, *_ = ask(lib, oracle_c, available, [], 61, CORE)
    assert rc == -8                                                            # EGPU_ERR_UNSAT: the scan said -1


def test_memory_resource_uses_the_memory_column(lib, oracle_c):
    available = ids(0, range(0, 300)) + ids(3, range(1000, 1200))
    rc, chosen, gpu, table, req = ask(lib, oracle_c, available, [], 150, MEM)
    assert rc == 0 and table == ([CORE_MAX] * 4, [300, 0, 0, 200]) and req == (1, 150)
    assert gpu == 3 and chosen == ids(3, range(1000, 1150))                    # 200 leaves 50, 300 leaves 150


def test_must_include_pins_the_gpu(lib, oracle_c):
    available = ids(0, range(40, 100)) + ids(2, range(75, 100)) + ids(1, range(0, 30))
    rc, chosen, gpu, table, req = ask(lib, oracle_c, available, ["0-55", "0-41"], 25, CORE)
    assert rc == 0 and table == ([60, 0, 0], [MEM_MAX, 0, 0])                  # the other GPUs are closed off
    assert gpu == 0 and chosen[:2] == ["0-55", "0-41"]
    assert chosen[2:] == [i for i in ids(0, range(40, 100)) if i not in ("0-55", "0-41")][:23]
    assert ask(lib, oracle_c, available, ["0-55", "1-03"], 25, CORE)[0] == -8  # must-include spans two GPUs
    assert ask(lib, oracle_c, available, ["2-80"], 26, CORE)[0] == -8           # pinned GPU is too small
    assert ask(lib, oracle_c, available, ["5-00"], 1, CORE)[0] == -8            # must-include not among the available


@pytest.mark.parametrize("bad", ["zz", "1-1", "01-05", "", "1-05x"])
def test_malformed_ids_fail_before_the_device_is_asked(lib, oracle_c, bad):
    rc, _, _, table, _ = ask(lib, oracle_c, ids(0, range(0, 10)) + [bad], [], 1, CORE)
    assert rc == -7 and table is None


def test_edges(lib, oracle_c):
    assert ask(lib, oracle_c, [], [], 0, CORE)[0] == 0                          # nothing asked, nothing available
    assert ask(lib, oracle_c, [], [], 1, CORE)[0] == -8
    assert ask(lib, oracle_c, ids(0, range(0, 100)) + ids(1, range(0, 100)), [], 101, CORE)[0] == -8  # > one card: not v1
    assert ask(lib, oracle_c, ["64-00"], [], 1, CORE)[0] == -1                  # beyond EGPU_MAX_DEVICES
    assert ask(lib, oracle_c, ids(0, range(0, 3)), ids(0, range(0, 3)), 2, CORE)[0] == -8  # more must-include than asked
