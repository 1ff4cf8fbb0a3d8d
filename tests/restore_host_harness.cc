// Test harness (CPU): links elastic-gpu-agent_b200/csrc/egpu_restore.cc - the host side of
// egpu_table_restore: JSON reader, symlink parsers, flattening - against a RECORDING stand-in for
// egpu_table_restore_flat, so the host logic is checked by `pytest -m "not gpu"` without a
// device.  Nothing here is a CPU implementation of the device work: the stand-in computes
// nothing, it only captures the arrays it is handed.  Built by tests/test_restore_host.py.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "egpu_restore.h"

namespace {
struct Capture {
    std::string flat, hash8, last_error;
    std::vector<int64_t> id_off, set_off, link_off;
    std::vector<int32_t> resource, link_gpu, cap;
    int32_t D = 0;
    int flags = 0;
    int calls = 0;
} g;
}  // namespace

void egpu_note_error(egpu_ctx*, const char* msg) { g.last_error = msg ? msg : ""; }

extern "C" {

int egpu_table_restore_flat(egpu_ctx*, const char* ids_flat, const int64_t* id_offsets, int64_t n_ids,
                            const int64_t* set_offsets, int64_t n_sets, const char* set_hash8, const int32_t* set_resource,
                            const int64_t* link_offsets, const int32_t* link_gpu, const int32_t* cap_core,
                            const int32_t* cap_mem, int32_t D, int flags, int32_t* out_table, int32_t* out_status) {
    g.calls += 1;
    g.id_off.assign(id_offsets, id_offsets + n_ids + 1);
    g.flat.assign(ids_flat ? ids_flat : "", static_cast<size_t>(id_offsets[n_ids]));
    g.set_off.assign(set_offsets, set_offsets + n_sets + 1);
    g.hash8.assign(set_hash8 ? set_hash8 : "", static_cast<size_t>(8 * n_sets));
    g.resource.assign(set_resource, set_resource + n_sets);
    g.link_off.assign(link_offsets, link_offsets + n_sets + 1);
    g.link_gpu.assign(link_gpu, link_gpu + link_offsets[n_sets]);
    g.cap.assign(cap_core, cap_core + D);
    g.cap.insert(g.cap.end(), cap_mem, cap_mem + D);
    g.D = D;
    g.flags = flags;
    for (int i = 0; i < 3 * D; ++i) out_table[i] = 0;
    for (int64_t q = 0; q < n_sets; ++q) out_status[q] = static_cast<int32_t>(q % EGPU_REC_STATUS_COUNT);  // a recognisable pattern
    return EGPU_OK;
}

// accessors for ctypes
int64_t cap_n(int what) {
    switch (what) {
        case 0: return static_cast<int64_t>(g.flat.size());
        case 1: return static_cast<int64_t>(g.id_off.size());
        case 2: return static_cast<int64_t>(g.set_off.size());
        case 3: return static_cast<int64_t>(g.hash8.size());
        case 4: return static_cast<int64_t>(g.resource.size());
        case 5: return static_cast<int64_t>(g.link_off.size());
        case 6: return static_cast<int64_t>(g.link_gpu.size());
        case 7: return g.calls;
        case 8: return g.flags;
        default: return -1;
    }
}
const void* cap_ptr(int what) {
    switch (what) {
        case 0: return g.flat.data();
        case 1: return g.id_off.data();
        case 2: return g.set_off.data();
        case 3: return g.hash8.data();
        case 4: return g.resource.data();
        case 5: return g.link_off.data();
        case 6: return g.link_gpu.data();
        default: return nullptr;
    }
}
const char* cap_error(void) { return g.last_error.c_str(); }

}  // extern "C"
