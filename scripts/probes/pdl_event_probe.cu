// Does an event recorded after a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization wait for that kernel to finish?
// Variants: attribute on/off; kernel calls griddepcontrol.launch_dependents early or never.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(volatile int* done, long long cycles, int trigger_early, int wait_first) {
    if (wait_first) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (threadIdx.x == 0 && blockIdx.x == 0) *done = 1;
}
int main() {
    int* d;  cudaMallocManaged(&d, sizeof(int));
    int* hflag; cudaMallocHost(&hflag, sizeof(int));
    cudaStream_t s; cudaStreamCreate(&s);
    for (int attr = 0; attr < 2; ++attr)
        for (int trig = 0; trig < 2; ++trig)
            for (int nk = 1; nk <= 2; ++nk) {
                *hflag = 0; cudaDeviceSynchronize();
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0, s);
                for (int k = 0; k < nk; ++k) {
                    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(8); cfg.blockDim = dim3(32); cfg.stream = s;
                    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    a[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = a; cfg.numAttrs = attr ? 1 : 0;
                    cudaLaunchKernelEx(&cfg, spin, (volatile int*)hflag, 20000000ll, trig, 1);  // ~10 ms
                }
                cudaEventRecord(e1, s);
                cudaEventSynchronize(e1);
                const int seen = *(volatile int*)hflag;
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                cudaDeviceSynchronize();
                printf("attr=%d early_trigger=%d kernels=%d: event elapsed %.2f ms, kernel finished when event completed: %s\n", attr, trig, nk, ms, seen ? "yes" : "NO");
            }
    return 0;
}
