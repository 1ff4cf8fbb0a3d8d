// Test harness (CPU): links elastic-gpu-agent_b200/csrc/egpu_plugin.cc - the host logic of
// GetPreferredAllocation - against a RECORDING stand-in for the device entry point it uses,
// egpu_bestfit_query.  The stand-in for the scan computes nothing: it hands
// back the answer the test scripted (which the test takes from the oracle) and records what it
// was asked.  Built by tests/test_plugin_host_cpu.py.
#include <cstdint>
#include <vector>

#include "egpu_plugin.h"

namespace {
std::vector<int32_t> g_fc, g_fm;
int32_t g_req_core = -1, g_req_mem = -1, g_answer = -1;
int g_sets = 0, g_scans = 0;
// optional: instead of the scripted answer, ask a callback the TEST installs (tests/test_kubelet_grpc.py
// passes the oracle's oracle_pick_one - test infrastructure answering for the device, never the product)
using pick_fn = int32_t (*)(const int32_t*, const int32_t*, int32_t, int32_t, int32_t);
pick_fn g_pick = nullptr;
}  // namespace

extern "C" {

int egpu_bestfit_query(egpu_ctx*, const int32_t* free_core, const int32_t* free_mem, int32_t D, const int32_t* req_core,
                       const int32_t* req_mem, int64_t R, int32_t* out_idx) {
    if (R != 1) return EGPU_ERR_INVALID;
    g_fc.assign(free_core, free_core + D);
    g_fm.assign(free_mem, free_mem + D);
    g_sets += 1;
    g_req_core = req_core[0];
    g_req_mem = req_mem[0];
    g_scans += 1;
    out_idx[0] = g_pick ? g_pick(free_core, free_mem, D, req_core[0], req_mem[0]) : g_answer;
    return EGPU_OK;
}

void stub_use_callback(pick_fn f) { g_pick = f; }

void stub_script_answer(int32_t idx) { g_answer = idx; }
int32_t stub_table(int32_t* fc, int32_t* fm) {
    for (size_t d = 0; d < g_fc.size(); ++d) {
        fc[d] = g_fc[d];
        fm[d] = g_fm[d];
    }
    return static_cast<int32_t>(g_fc.size());
}
void stub_request(int32_t* core, int32_t* mem, int32_t* sets, int32_t* scans) {
    *core = g_req_core;
    *mem = g_req_mem;
    *sets = g_sets;
    *scans = g_scans;
}

}  // extern "C"
