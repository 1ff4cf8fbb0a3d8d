/*
 * oracle/bestfit_oracle.c — CPU restatement of the best-fit allocation spec.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * elastic-gpu-agent_b200/, its C-ABI library) may include, link or call this
 * file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker and as the timed CPU baseline.
 *
 * PARITY UNPINNED.  The reference (elastic-ai/elastic-gpu-agent @ 2609107)
 * contains no best-fit scoring loop to restate: its Allocate hashes device-ID
 * strings (pkg/plugins/gpushare.go:36-86, :171-211) and the chosen GPU index
 * arrives in a pod annotation written by a different product
 * (pkg/plugins/gpushare.go:107-125, :232-245).  GetPreferredAllocation, the
 * kubelet hook where a node-local choice belongs, is an empty stub
 * (pkg/plugins/base.go:94-96).  What this file follows is therefore the
 * builder-defined specification,
 * constrained by the units the reference does pin:
 *   - core is an integer percentage, 100 per card
 *     (pkg/common/const.go:4, pkg/plugins/gpushare.go:24-33)
 *   - memory is an integer number of MiB
 *     (pkg/plugins/gpushare.go:159-168, pkg/operator/base.go:35-39)
 *   - a device index is the NVML order = N of /dev/nvidiaN
 *     (pkg/operator/base.go:29-33, pkg/operator/gpushare.go:10,32)
 *   - commits are serialised (pkg/plugins/gpushare.go:114,239) -> sequential mode
 * The known-answer vectors it is pinned to are hand-derived (tests/golden/),
 * and it is cross-checked against an independent numpy restatement
 * (oracle/bestfit_np.py).
 *
 * Written as the obvious scalar loop on purpose: per request, walk every
 * device, compare, keep the smallest (leftover core, leftover mem, index).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_DEVICES 64
#define ORACLE_CORE_MAX 100
#define ORACLE_MEM_MAX ((1 << 18) - 1)

/* 0 when the table is inside the spec's domain */
int oracle_table_valid(const int32_t* free_core, const int32_t* free_mem, int32_t D) {
    if (!free_core || !free_mem || D < 1 || D > ORACLE_MAX_DEVICES) return -1;
    for (int32_t d = 0; d < D; ++d) {
        if (free_core[d] < 0 || free_core[d] > ORACLE_CORE_MAX) return -1;
        if (free_mem[d] < 0 || free_mem[d] > ORACLE_MEM_MAX) return -1;
    }
    return 0;
}

/* spec §2.3: index of the best-fit device or -1 */
static inline int32_t oracle_pick(const int32_t* free_core, const int32_t* free_mem,
                                  int32_t D, int32_t core, int32_t mem) {
    int32_t best = -1;
    int32_t best_lc = 0, best_lm = 0;
    if (core < 0 || mem < 0) return -1;
    for (int32_t d = 0; d < D; ++d) {
        if (free_core[d] < core || free_mem[d] < mem) continue;
        int32_t lc = free_core[d] - core;
        int32_t lm = free_mem[d] - mem;
        /* strict "<" in index order keeps the lowest index on ties */
        if (best < 0 || lc < best_lc || (lc == best_lc && lm < best_lm)) {
            best = d;
            best_lc = lc;
            best_lm = lm;
        }
    }
    return best;
}

int32_t oracle_pick_one(const int32_t* free_core, const int32_t* free_mem, int32_t D,
                        int32_t core, int32_t mem) {
    return oracle_pick(free_core, free_mem, D, core, mem);
}

static int32_t sat_i32(int64_t v) {
    if (v > INT32_MAX) return INT32_MAX;
    if (v < INT32_MIN) return INT32_MIN;
    return (int32_t)v;
}

/* spec §2.4 table' = table - delta, saturated, with oversubscription flags.
 * table_out is int32[3*D]: free_core', free_mem', oversub. */
void oracle_apply_delta(const int32_t* free_core, const int32_t* free_mem, int32_t D,
                        const int64_t* delta_core, const int64_t* delta_mem,
                        int32_t* table_out) {
    for (int32_t d = 0; d < D; ++d) {
        int64_t c = (int64_t)free_core[d] - delta_core[d];
        int64_t m = (int64_t)free_mem[d] - delta_mem[d];
        table_out[d] = sat_i32(c);
        table_out[D + d] = sat_i32(m);
        table_out[2 * D + d] = (c < 0 || m < 0) ? 1 : 0;
    }
}

/* Snapshot mode (spec §2.4): every request against the same table.
 * nthreads <= 1: single scalar thread; otherwise OpenMP over request rows. */
int oracle_bestfit_snapshot(const int32_t* free_core, const int32_t* free_mem, int32_t D,
                            const int32_t* req_core, const int32_t* req_mem, int64_t R,
                            int32_t* out_idx, int64_t* delta_core, int64_t* delta_mem,
                            int32_t* table_out, int nthreads) {
    if (oracle_table_valid(free_core, free_mem, D) != 0 || R < 0) return -1;
    int64_t dc[ORACLE_MAX_DEVICES], dm[ORACLE_MAX_DEVICES];
    memset(dc, 0, sizeof dc);
    memset(dm, 0, sizeof dm);
    if (nthreads <= 1) {
        for (int64_t r = 0; r < R; ++r) {
            int32_t i = oracle_pick(free_core, free_mem, D, req_core[r], req_mem[r]);
            out_idx[r] = i;
            if (i >= 0) {
                dc[i] += req_core[r];
                dm[i] += req_mem[r];
            }
        }
    } else {
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
        {
            int64_t lc[ORACLE_MAX_DEVICES], lm[ORACLE_MAX_DEVICES];
            memset(lc, 0, sizeof lc);
            memset(lm, 0, sizeof lm);
#pragma omp for schedule(static)
            for (int64_t r = 0; r < R; ++r) {
                int32_t i = oracle_pick(free_core, free_mem, D, req_core[r], req_mem[r]);
                out_idx[r] = i;
                if (i >= 0) {
                    lc[i] += req_core[r];
                    lm[i] += req_mem[r];
                }
            }
#pragma omp critical
            for (int32_t d = 0; d < D; ++d) {
                dc[d] += lc[d];
                dm[d] += lm[d];
            }
        }
#else
        return -2;
#endif
    }
    if (delta_core) memcpy(delta_core, dc, sizeof(int64_t) * (size_t)D);
    if (delta_mem) memcpy(delta_mem, dm, sizeof(int64_t) * (size_t)D);
    if (table_out) oracle_apply_delta(free_core, free_mem, D, dc, dm, table_out);
    return 0;
}

/* Prefix-commit (spec §2.5): the snapshot choice, but request r
 * commits only while the running demand of its device - over ALL rows <= r that chose it,
 * committed or not - still fits free[d]; otherwise out_idx[r] = -2 (DEFERRED).  The demand
 * sums and table' count committed rows only, so table' is never negative. */
int oracle_bestfit_prefix_commit(const int32_t* free_core, const int32_t* free_mem, int32_t D,
                                 const int32_t* req_core, const int32_t* req_mem, int64_t R,
                                 int32_t* out_idx, int64_t* delta_core, int64_t* delta_mem,
                                 int32_t* table_out) {
    if (oracle_table_valid(free_core, free_mem, D) != 0 || R < 0) return -1;
    int64_t run_c[ORACLE_MAX_DEVICES], run_m[ORACLE_MAX_DEVICES], dc[ORACLE_MAX_DEVICES], dm[ORACLE_MAX_DEVICES];
    memset(run_c, 0, sizeof run_c);
    memset(run_m, 0, sizeof run_m);
    memset(dc, 0, sizeof dc);
    memset(dm, 0, sizeof dm);
    for (int64_t r = 0; r < R; ++r) {
        int32_t d = oracle_pick(free_core, free_mem, D, req_core[r], req_mem[r]);
        if (d >= 0) {
            run_c[d] += req_core[r];
            run_m[d] += req_mem[r];
            if (run_c[d] <= free_core[d] && run_m[d] <= free_mem[d]) {
                dc[d] += req_core[r];
                dm[d] += req_mem[r];
            } else {
                d = -2;
            }
        }
        out_idx[r] = d;
    }
    if (delta_core) memcpy(delta_core, dc, sizeof(int64_t) * (size_t)D);
    if (delta_mem) memcpy(delta_mem, dm, sizeof(int64_t) * (size_t)D);
    if (table_out) oracle_apply_delta(free_core, free_mem, D, dc, dm, table_out);
    return 0;
}

/* Rounds (spec §2.5, "multi-round deferred retry"): committing
 * prefix-commit rounds; each round re-submits the rows the previous one deferred, in order,
 * against the committed table.  free_core/free_mem are updated in place.  Returns the number
 * of rounds run (>= 0) or a negative error; *left = rows still deferred. */
int oracle_bestfit_rounds(int32_t* free_core, int32_t* free_mem, int32_t D, const int32_t* req_core,
                          const int32_t* req_mem, int64_t R, int32_t* out_idx, int64_t* delta_core,
                          int64_t* delta_mem, int32_t max_rounds, int64_t* left) {
    if (oracle_table_valid(free_core, free_mem, D) != 0 || R < 0 || max_rounds < 1) return -1;
    int64_t tc[ORACLE_MAX_DEVICES], tm[ORACLE_MAX_DEVICES];
    memset(tc, 0, sizeof tc);
    memset(tm, 0, sizeof tm);
    int64_t* rows = (int64_t*)malloc(sizeof(int64_t) * (size_t)(R > 0 ? R : 1));
    if (!rows) return -3;
    int64_t n = R;
    for (int64_t r = 0; r < R; ++r) rows[r] = r;
    int rounds = 0;
    while (n > 0 && rounds < max_rounds) {
        int64_t run_c[ORACLE_MAX_DEVICES], run_m[ORACLE_MAX_DEVICES], dc[ORACLE_MAX_DEVICES], dm[ORACLE_MAX_DEVICES];
        memset(run_c, 0, sizeof run_c);
        memset(run_m, 0, sizeof run_m);
        memset(dc, 0, sizeof dc);
        memset(dm, 0, sizeof dm);
        int64_t kept = 0;
        for (int64_t i = 0; i < n; ++i) {
            const int64_t r = rows[i];
            int32_t d = oracle_pick(free_core, free_mem, D, req_core[r], req_mem[r]);  /* table of the round's start */
            if (d >= 0) {
                run_c[d] += req_core[r];
                run_m[d] += req_mem[r];
                if (run_c[d] <= free_core[d] && run_m[d] <= free_mem[d]) {
                    dc[d] += req_core[r];
                    dm[d] += req_mem[r];
                } else {
                    d = -2;
                    rows[kept++] = r;
                }
            }
            out_idx[r] = d;
        }
        for (int32_t d = 0; d < D; ++d) {  /* commit the round */
            free_core[d] -= (int32_t)dc[d];
            free_mem[d] -= (int32_t)dm[d];
            tc[d] += dc[d];
            tm[d] += dm[d];
        }
        n = kept;
        ++rounds;
    }
    free(rows);
    if (delta_core) memcpy(delta_core, tc, sizeof(int64_t) * (size_t)D);
    if (delta_mem) memcpy(delta_mem, tm, sizeof(int64_t) * (size_t)D);
    if (left) *left = n;
    return rounds;
}

/* Sequential mode (spec §2.6).  free_core/free_mem are updated in place.
 * kind 0 = ALLOC(core=a, mem=b); kind 1 = FREE(event index a). */
int oracle_replay(int32_t* free_core, int32_t* free_mem, int32_t D, const int32_t* kind,
                  const int32_t* a, const int32_t* b, int64_t E, int32_t* out_idx) {
    if (oracle_table_valid(free_core, free_mem, D) != 0 || E < 0) return -1;
    /* live[i] = device an ALLOC event currently holds, -1 otherwise */
    int8_t* live = (int8_t*)malloc((size_t)(E > 0 ? E : 1));
    if (!live) return -3;
    memset(live, -1, (size_t)(E > 0 ? E : 1));
    for (int64_t i = 0; i < E; ++i) {
        if (kind[i] == 0) {
            int32_t d = oracle_pick(free_core, free_mem, D, a[i], b[i]);
            out_idx[i] = d;
            if (d >= 0) {
                free_core[d] -= a[i];
                free_mem[d] -= b[i];
                live[i] = (int8_t)d;
            }
        } else {
            int64_t t = a[i];
            int32_t d = -1;
            if (kind[i] == 1 && t >= 0 && t < i && kind[t] == 0 && live[t] >= 0) {
                d = live[t];
                free_core[d] += a[t];
                free_mem[d] += b[t];
                live[t] = -1;
            }
            out_idx[i] = d;
        }
    }
    free(live);
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
