/*
 * egpu_plugin.h — C ABI of the plugin-side host logic around the best-fit path:
 * the device-ID codec and GetPreferredAllocation for one container request.
 *
 * Reference interfaces (elastic-ai/elastic-gpu-agent @ 2609107):
 *   egpu_device_id_format / egpu_device_id_parse
 *       the fake device IDs the plugins advertise, fmt.Sprintf("%d-%02d", gpu, j)
 *       (pkg/plugins/gpushare.go:28 for gpu-core, :163 for gpu-memory);
 *   egpu_preferred_allocation
 *       pluginapi.ContainerPreferredAllocationRequest -> ...Response
 *       (vendor/k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.pb.go:564-571, 675-677;
 *       api.proto:133-150), the body baseDevicePlugin.GetPreferredAllocation leaves empty
 *       (pkg/plugins/base.go:94-96).
 *
 * The choice itself is made by the CUDA best-fit scan (egpu_bestfit_query); this file
 * only turns ID strings into the capacity table and the answer back into IDs.  No CPU
 * fallback: without a context (no GPU) the call fails.
 */
#ifndef EGPU_PLUGIN_H
#define EGPU_PLUGIN_H

#include <stdint.h>

#include "egpu_alloc.h"

#ifdef __cplusplus
extern "C" {
#endif

#define EGPU_RESOURCE_CORE 0  /* elasticgpu.io/gpu-core:   one ID per percent, 100 per GPU */
#define EGPU_RESOURCE_MEM  1  /* elasticgpu.io/gpu-memory: one ID per MiB */

/* "%d-%02d".  Returns the length written (excluding NUL) or EGPU_ERR_INVALID if it
 * does not fit in cap bytes or an argument is negative. */
int egpu_device_id_format(int32_t gpu, int64_t unit, char* out, int64_t cap);

/* Strict inverse: decimal gpu, '-', decimal unit of at least two digits; nothing else.
 * EGPU_ERR_PARSE otherwise. */
int egpu_device_id_parse(const char* id, int32_t* gpu, int64_t* unit);

/* One ContainerPreferredAllocationRequest.
 *   available_ids[n_available], must_include_ids[n_must]: NUL-terminated ID strings
 *   allocation_size: number of IDs (= units of `resource`) the container asks for
 * The free table is what the available IDs say: per GPU, the number of available units
 * of this resource (the other resource is unconstrained).  The request is
 * allocation_size units; the GPU is the best fit (tightest leftover, lowest index).
 * must_include IDs pin the GPU: they must all sit on one GPU that can hold the request.
 * Output: out_positions[allocation_size] = indices into available_ids of the chosen IDs,
 * must-include IDs first, then the lowest unit numbers of the chosen GPU; *out_gpu = the
 * GPU index.  EGPU_ERR_UNSAT when no single GPU can satisfy the request (kubelet then
 * falls back to its own choice), EGPU_ERR_PARSE on a malformed ID.
 * Whole-card requests (gpu-core, allocation_size > 100): the reference hands such a container
 * allocation_size / 100 whole GPUs (pkg/plugins/gpushare.go:62-69).  The size must be a multiple of
 * 100; the GPUs are chosen one after the other, each the best fit among the cards that are still
 * completely available (must-include GPUs first); the answer lists all 100 IDs of every chosen card,
 * *out_gpu is the first one.
 * The context's own capacity table is not touched (the availability table of the request is
 * scored through egpu_bestfit_query), so the context that tracks the node's committed
 * placement can serve these calls too, concurrently with commits and replays. */
int egpu_preferred_allocation(egpu_ctx* ctx, const char* const* available_ids, int64_t n_available,
                              const char* const* must_include_ids, int64_t n_must,
                              int32_t allocation_size, int resource, int32_t* out_positions,
                              int32_t* out_gpu);

#ifdef __cplusplus
}
#endif
#endif /* EGPU_PLUGIN_H */
