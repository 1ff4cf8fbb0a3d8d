"""Writes tests/golden/restore_records.json: a hand-derived known-answer scenario for
placement-state restore (include/egpu_restore.h) in the reference's stored formats
(pkg/types/pod.go:39-58, pkg/operator/gpushare.go:31-55).  The expected table and statuses
below are written out by hand; the script asserts that the CPU oracle reproduces them before
it writes the file.  Run from the repo root: python tests/golden/make_restore_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import restore_py as R  # noqa: E402


def ids(gpu, lo, hi):
    return ["%d-%02d" % (gpu, j) for j in range(lo, hi)]


def main():
    D = 4
    cap_core, cap_mem = [100] * D, [1000] * D
    records, links = [], []

    def link(id_list, ordinal, gpu, bare=False, hash_override=None):
        h = hash_override or R.device_hash(id_list)
        links.append([("" if bare else "elastic-gpu-") + "%s-%d" % (h, ordinal), "/dev/nvidia%d" % gpu])
        links.append(["elastic-gpuctl-%s-%d" % (h, ordinal), "/dev/nvidiactl"])  # the ctl twin: skipped

    a = ids(0, 0, 25)                                   # 25 % on GPU 1
    records.append(R.marshal_record("default", "a", {"main": (a, R.CORE)}))
    link(a, 0, 1)
    b_mem, b_core = ids(2, 0, 300), ids(3, 0, 10)       # 300 MiB on GPU 1, 10 % on GPU 0
    records.append(R.marshal_record("default", "b", {"main": (b_mem, R.MEM), "side": (b_core, R.CORE)}))
    link(b_mem, 0, 1)
    link(b_core, 0, 0)
    c = ids(0, 0, 100) + ids(1, 0, 100)                 # two whole cards: GPUs 2 and 3
    records.append(R.marshal_record("default", "c", {"main": (c, R.CORE)}))
    link(c, 0, 2)
    link(c, 1, 3)
    d = ids(1, 40, 45)                                  # never started: no symlink
    records.append(R.marshal_record("default", "d", {"main": (d, R.CORE)}))
    e = ids(2, 50, 55)                                  # stored hash does not match the list
    k, v = R.marshal_record("default", "e", {"main": (e, R.CORE)})
    v = v.replace(R.device_hash(e).encode(), b"deadbeef")
    records.append((k, v))
    link(e, 0, 0, hash_override="deadbeef")
    # another plugin's resource, a nil entry, an empty list
    records.append((b"kube-system/f",
                    b'{"main":{"Hash":"ba7816bf","List":["a","b","c"],"ResourceName":"tke.cloud.tencent.com/qgpu-core"},'
                    b'"nil":null,"none":{"Hash":"e3b0c442","List":[],"ResourceName":"elasticgpu.io/gpu-core"}}'))
    g = ids(3, 5, 95)                                   # 90 % more on GPU 1: 115 > 100
    records.append(R.marshal_record("default", "g", {"main": (g, R.CORE)}))
    link(g, 0, 1)
    h = ids(1, 0, 10)                                   # link points outside the 4-GPU table
    records.append(R.marshal_record("default", "h", {"main": (h, R.MEM)}))
    link(h, 0, 7)
    i = ids(3, 97, 100)                                 # encoding/json leniency; bare link name
    records.append((b"default/i",
                    ('{ "we\\u0069rd\\"name" : { "resourcename" : "elasticgpu.io/gpu-core", "extra": [1, {"x": null}, "y"],\n'
                     '  "LIST" : ["3-97", "3-98", "3-99"], "hash" : "%s" } }' % R.device_hash(i)).encode()))
    link(i, 0, 0, bare=True)
    links.append(["nvidia0", "/dev/nvidia0"])           # unrelated directory entries
    links.append(["elastic-gpu-zzz", "/dev/nvidia1"])

    expect = {
        "verify": {"free_core": [87, 0, 0, 0], "free_mem": [1000, 700, 1000, 1000], "oversub": [0, 1, 0, 0],
                   "counts": [6, 2, 1, 2, 1], "record_status": [0, 0, 0, 3, 4, 2, 0, 3, 0]},
        "no_verify": {"free_core": [82, 0, 0, 0], "free_mem": [1000, 700, 1000, 1000], "oversub": [0, 1, 0, 0],
                      "counts": [7, 2, 1, 2, 0], "record_status": [0, 0, 0, 3, 0, 2, 0, 3, 0]},
    }
    for mode, verify in (("verify", True), ("no_verify", False)):
        fc, fm, ov, counts, rstat = R.restore(records, links, cap_core, cap_mem, verify=verify)
        got = {"free_core": fc, "free_mem": fm, "oversub": ov, "counts": counts, "record_status": rstat}
        assert got == expect[mode], (mode, got, expect[mode])
    bad = [["nokey", "{}"], ["a/b/c", "{}"], ["a/b", "{"], ["a/b", "[1]"], ["a/b", '{"c":{"List":[1]}}'],
           ["a/b", '{"c":{"Hash":"00000000","List":["x-1"],"ResourceName":"elasticgpu.io/gpu-core"}}']]
    for key, val in bad:
        try:
            R.restore([(key.encode(), val.encode())], [], cap_core, cap_mem)
        except R.RestoreParseError:
            continue
        raise AssertionError(("oracle accepted", key, val))
    out = {"comment": "hand-derived; see make_restore_golden.py", "D": D, "cap_core": cap_core, "cap_mem": cap_mem,
           "records": [[k.decode(), v.decode()] for k, v in records], "links": links, "expect": expect,
           "parse_errors": bad}
    with open(os.path.join(ROOT, "tests", "golden", "restore_records.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote restore_records.json:", len(records), "records,", len(links), "links")


if __name__ == "__main__":
    main()
