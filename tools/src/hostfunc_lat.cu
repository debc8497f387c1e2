// hostfunc_lat.cu -- what a cudaLaunchHostFunc in the middle of a stream costs on this box: the gap between the
// kernel before and the kernel after an empty host function, and whether host functions of two streams overlap.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/bin/hostfunc_lat tools/src/hostfunc_lat.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <thread>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__global__ void tiny(int* p) { if (threadIdx.x == 0) atomicAdd(p, 1); }
static void CUDART_CB noop(void*) {}
static void CUDART_CB busy(void* us) {
    const auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() < *(double*)us) {}
}
int main() {
    int* d; CK(cudaMalloc(&d, 4)); CK(cudaMemset(d, 0, 4));
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t a, b, c, e2; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreate(&c)); CK(cudaEventCreate(&e2));
    for (int rep = 0; rep < 5; ++rep) {
        float ms0, ms1;
        CK(cudaEventRecord(a, s1)); tiny<<<1, 32, 0, s1>>>(d); tiny<<<1, 32, 0, s1>>>(d); CK(cudaEventRecord(b, s1));
        CK(cudaStreamSynchronize(s1)); CK(cudaEventElapsedTime(&ms0, a, b));
        CK(cudaEventRecord(a, s1)); tiny<<<1, 32, 0, s1>>>(d); CK(cudaLaunchHostFunc(s1, noop, nullptr)); tiny<<<1, 32, 0, s1>>>(d); CK(cudaEventRecord(b, s1));
        CK(cudaStreamSynchronize(s1)); CK(cudaEventElapsedTime(&ms1, a, b));
        printf("two kernels: %.1f us; with an empty host function between them: %.1f us\n", ms0 * 1e3, ms1 * 1e3);
    }
    // ten in a row (steady state, host thread awake)
    {
        float ms;
        CK(cudaEventRecord(a, s1));
        for (int i = 0; i < 10; ++i) { tiny<<<1, 32, 0, s1>>>(d); CK(cudaLaunchHostFunc(s1, noop, nullptr)); }
        CK(cudaEventRecord(b, s1)); CK(cudaStreamSynchronize(s1)); CK(cudaEventElapsedTime(&ms, a, b));
        printf("10 x (kernel, empty host function) on one stream: %.1f us each\n", ms * 100);
    }
    // two streams, a 200 us host function each: overlapped or serialised?
    {
        double us = 200.0; float ms;
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, s1)); CK(cudaStreamWaitEvent(s2, a, 0));
        CK(cudaLaunchHostFunc(s1, busy, &us)); CK(cudaLaunchHostFunc(s2, busy, &us));
        CK(cudaEventRecord(e2, s2)); CK(cudaStreamWaitEvent(s1, e2, 0)); CK(cudaEventRecord(b, s1));
        CK(cudaStreamSynchronize(s1)); CK(cudaEventElapsedTime(&ms, a, b));
        printf("two streams x one 200 us host function: %.1f us in total (%s)\n", ms * 1e3, ms * 1e3 > 350 ? "serialised" : "overlapped");
    }
    return 0;
}
