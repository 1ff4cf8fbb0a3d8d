/*
 * A plain C client of the C ABI, linked against libegpu_alloc.so the way the cgo shim of
 * INTEGRATION.md would be: no Python, no C++, no CUDA headers.  Runs the cfg1 known-answer
 * vector, a GetPreferredAllocation call, a device-set hash, the
 * rounds form of the batch and a restore from one stored record.
 * Built and run by tests/test_c_client.py.  Exit code 0 = all checks passed.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "egpu_alloc.h"
#include "egpu_devhash.h"
#include "egpu_plugin.h"
#include "egpu_restore.h"

#define CHECK(cond, msg)                                  \
    do {                                                  \
        if (!(cond)) {                                    \
            fprintf(stderr, "FAILED: %s\n", msg);         \
            return 1;                                     \
        }                                                 \
    } while (0)

int main(void) {
    egpu_ctx* ctx = NULL;
    int rc = egpu_ctx_create(0, &ctx);
    if (rc == EGPU_ERR_NO_DEVICE) {
        printf("no CUDA device: %s\n", egpu_strerror(rc));
        return 77; /* the CPU-side test accepts this: there is no fallback to exercise */
    }
    CHECK(rc == EGPU_OK, "egpu_ctx_create");
    int32_t fc[8], fm[8];
    for (int d = 0; d < 8; ++d) { fc[d] = 100; fm[d] = 183359; }
    CHECK(egpu_table_set(ctx, fc, fm, 8) == EGPU_OK, "egpu_table_set");

    /* cfg1, snapshot: five identical requests all pick device 0 and oversubscribe it */
    int32_t core[5] = {25, 25, 25, 25, 25}, mem[5] = {1024, 1024, 1024, 1024, 1024}, idx[5];
    int64_t dc[8], dm[8];
    CHECK(egpu_bestfit_batch(ctx, core, mem, 5, idx, dc, dm, 0) == EGPU_OK, "egpu_bestfit_batch");
    for (int i = 0; i < 5; ++i) CHECK(idx[i] == 0, "snapshot idx");
    CHECK(dc[0] == 125 && dm[0] == 5120, "snapshot demand sums");

    /* the same with prefix-commit: the fifth is deferred, the table ends at exactly 0 core */
    CHECK(egpu_bestfit_batch(ctx, core, mem, 5, idx, dc, dm, EGPU_F_COMMIT | EGPU_F_PREFIX_COMMIT) == EGPU_OK, "prefix commit");
    CHECK(idx[3] == 0 && idx[4] == EGPU_IDX_DEFERRED && dc[0] == 100, "prefix-commit result");
    int32_t ov[8];
    CHECK(egpu_table_get(ctx, fc, fm, ov) == EGPU_OK && fc[0] == 0 && fm[0] == 183359 - 4096 && ov[0] == 0, "table after commit");

    /* cfg1, sequential: the deferred request now goes to device 1 */
    int32_t kind[1] = {EGPU_EV_ALLOC}, a[1] = {25}, b[1] = {1024}, out[1];
    CHECK(egpu_replay(ctx, kind, a, b, 1, out) == EGPU_OK && out[0] == 1, "egpu_replay");

    /* GetPreferredAllocation: 30 core units of GPU 0 and 25 of GPU 2 available, 25 asked */
    char ids[55][8];
    const char* avail[55];
    int n = 0;
    for (int j = 0; j < 30; ++j) { egpu_device_id_format(0, 40 + j, ids[n], 8); avail[n] = ids[n]; ++n; }
    for (int j = 0; j < 25; ++j) { egpu_device_id_format(2, 75 + j, ids[n], 8); avail[n] = ids[n]; ++n; }
    int32_t pos[25], gpu = -1;
    CHECK(egpu_preferred_allocation(ctx, avail, n, NULL, 0, 25, EGPU_RESOURCE_CORE, pos, &gpu) == EGPU_OK, "preferred allocation");
    CHECK(gpu == 2 && strcmp(avail[pos[0]], "2-75") == 0 && strcmp(avail[pos[24]], "2-99") == 0, "preferred allocation picks the exact fit");

    /* ... and it left the context's own table alone (it scores through the stateless egpu_bestfit_query) */
    CHECK(egpu_table_size(ctx) == 8, "the tracked table survives GetPreferredAllocation");
    CHECK(egpu_table_get(ctx, fc, fm, ov) == EGPU_OK && fc[0] == 0 && fc[1] == 75, "tracked table unchanged");
    /* the query itself: a what-if table of two GPUs; (30, 1) fits only the second */
    int32_t qfc[2] = {20, 40}, qfm[2] = {500, 500}, qc[2] = {30, 10}, qm[2] = {1, 1}, qi[2] = {-9, -9};
    CHECK(egpu_bestfit_query(ctx, qfc, qfm, 2, qc, qm, 2, qi) == EGPU_OK && qi[0] == 1 && qi[1] == 0, "egpu_bestfit_query");
    /* whole cards: 200 gpu-core units = two completely available GPUs (the reference's len/100) */
    {
        static char wid[300][8];
        static const char* wav[300];
        int wn = 0;
        for (int g = 0; g < 3; ++g)
            for (int j = 0; j < 100; ++j) { egpu_device_id_format(g, j, wid[wn], 8); wav[wn] = wid[wn]; ++wn; }
        static int32_t wpos[200];
        int32_t wg = -1;
        CHECK(egpu_preferred_allocation(ctx, wav, wn - 50, NULL, 0, 200, EGPU_RESOURCE_CORE, wpos, &wg) == EGPU_OK, "whole-card request");
        CHECK(wg == 0 && strcmp(wav[wpos[0]], "0-00") == 0 && strcmp(wav[wpos[199]], "1-99") == 0, "whole cards 0 and 1 (card 2 is half taken)");
        CHECK(egpu_preferred_allocation(ctx, wav, wn - 50, NULL, 0, 300, EGPU_RESOURCE_CORE, wpos, &wg) == EGPU_ERR_UNSAT, "three whole cards are not there");
    }
    /* caller-owned memory pinned in place: the same call, no staging copies */
    {
        static int32_t rcore[4096], rmem[4096], ridx[4096];
        for (int i = 0; i < 4096; ++i) { rcore[i] = 1 + i % 50; rmem[i] = 1 + i; ridx[i] = -9; }
        CHECK(egpu_host_register(ctx, rcore, sizeof rcore) == EGPU_OK && egpu_host_register(ctx, rmem, sizeof rmem) == EGPU_OK &&
              egpu_host_register(ctx, ridx, sizeof ridx) == EGPU_OK, "egpu_host_register");
        CHECK(egpu_bestfit_batch(ctx, rcore, rmem, 4096, ridx, dc, dm, 0) == EGPU_OK && ridx[0] >= 0 && ridx[4095] >= 0, "batch on registered memory");
        CHECK(egpu_host_unregister(ctx, rcore) == EGPU_OK && egpu_host_unregister(ctx, rmem) == EGPU_OK &&
              egpu_host_unregister(ctx, ridx) == EGPU_OK, "egpu_host_unregister");
    }

    /* types.NewDevice([...]).Hash of {"3-07"} = first 8 hex digits of sha256("3-07") */
    const char* one[1] = {"3-07"};
    char h[9];
    CHECK(egpu_device_hash(ctx, one, 1, h) == EGPU_OK, "egpu_device_hash");
    printf("hash(3-07) = %s, launches = %lld\n", h, (long long)egpu_launch_count(ctx));

    /* rounds: the cfg1 batch to its fixed point is the sequential answer, in two rounds */
    for (int d = 0; d < 8; ++d) { fc[d] = 100; fm[d] = 183359; }
    CHECK(egpu_table_set(ctx, fc, fm, 8) == EGPU_OK, "egpu_table_set (2)");
    int32_t rounds = 0;
    int64_t left = -1;
    CHECK(egpu_bestfit_batch_rounds(ctx, core, mem, 5, idx, dc, dm, 16, &rounds, &left) == EGPU_OK, "egpu_bestfit_batch_rounds");
    CHECK(idx[3] == 0 && idx[4] == 1 && rounds == 2 && left == 0 && dc[0] == 100 && dc[1] == 25, "rounds result");

    /* restore: one stored record (pkg/types/pod.go:55-58) + its symlink -> 3 % less on GPU 2 */
    const char* keys[1] = {"default/pod"};
    const char* vals[1] = {"{\"main\":{\"Hash\":\"e4b78a7c\",\"List\":[\"0-00\",\"0-01\",\"3-07\"],"
                           "\"ResourceName\":\"elasticgpu.io/gpu-core\"}}"};
    int64_t klen[1] = {(int64_t)strlen(keys[0])}, vlen[1] = {(int64_t)strlen(vals[0])};
    const char* lnames[2] = {"elastic-gpu-e4b78a7c-0", "elastic-gpuctl-e4b78a7c-0"};
    const char* ltargets[2] = {"/dev/nvidia2", "/dev/nvidiactl"};
    int32_t capc[4] = {100, 100, 100, 100}, capm[4] = {1000, 1000, 1000, 1000}, table[12], rstat[1];
    int64_t counts[EGPU_REC_STATUS_COUNT];
    CHECK(egpu_table_restore(ctx, keys, klen, vals, vlen, 1, lnames, ltargets, 2, capc, capm, 4,
                             EGPU_RESTORE_VERIFY | EGPU_RESTORE_INSTALL, table, counts, rstat) == EGPU_OK, "egpu_table_restore");
    CHECK(table[2] == 97 && table[0] == 100 && table[4 + 2] == 1000 && counts[EGPU_REC_OK] == 1 && rstat[0] == EGPU_REC_OK, "restore result");
    CHECK(egpu_table_size(ctx) == 4, "restored table installed");
    egpu_ctx_destroy(ctx);
    printf("c abi client ok\n");
    return 0;
}
