// pcie_gather.cu -- an SM-issued gather of 16-byte rows from page-locked host memory, three ways, alone and with
// the step's DMA copies (2.2 MB H2D + 2.9 MB D2H) running beside it: is a device-side gather of the candidate
// rows (640 x 64 per C2 step) a substitute for the host-side gather?
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/bin/pcie_gather tools/src/pcie_gather.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// (a) one 16-byte load per row
__global__ void gather16(const float4* __restrict__ host, const int* __restrict__ idx, float4* __restrict__ out, int rows) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < rows) out[t] = host[idx[t]];
}
// (b) lane pairs load the two halves of the row's 32-byte sector (one sector request per pair), the right half is kept
__global__ void gather32pair(const float4* __restrict__ host, const int* __restrict__ idx, float4* __restrict__ out, int rows) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = t >> 1, half = t & 1;
    const int i = r < rows ? idx[r] : 0;
    float4 v = host[(i & ~1) + half];
    const float4 o = make_float4(__shfl_xor_sync(0xffffffffu, v.x, 1), __shfl_xor_sync(0xffffffffu, v.y, 1),
                                 __shfl_xor_sync(0xffffffffu, v.z, 1), __shfl_xor_sync(0xffffffffu, v.w, 1));
    if (r < rows && half == 0) out[r] = (i & 1) ? o : v;
}
// (c) one thread loads the whole 32-byte sector with two 16-byte loads
__global__ void gather32one(const float4* __restrict__ host, const int* __restrict__ idx, float4* __restrict__ out, int rows) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows) return;
    const int i = idx[t];
    const float4 a = host[i & ~1], b = host[(i & ~1) + 1];
    out[t] = (i & 1) ? b : a;
}

int main() {
    const int table = 64 * 8649, rows = 64 * 640;
    float4* host; CK(cudaHostAlloc(&host, (size_t)table * 16, cudaHostAllocDefault)); memset(host, 1, (size_t)table * 16);
    std::vector<int> hidx(rows);
    unsigned s = 777u;
    for (int b = 0; b < 64; ++b) for (int r = 0; r < 640; ++r) { s = s * 1664525u + 1013904223u; hidx[b * 640 + r] = b * 8649 + (int)((s >> 8) % 8649u); }
    int* idx; float4* out; CK(cudaMalloc(&idx, rows * 4)); CK(cudaMalloc(&out, (size_t)rows * 16));
    CK(cudaMemcpy(idx, hidx.data(), rows * 4, cudaMemcpyHostToDevice));
    char *hin, *hout, *din, *dout;
    const size_t nin = 2280000, nout = 2900000;
    CK(cudaHostAlloc(&hin, nin, cudaHostAllocDefault)); CK(cudaHostAlloc(&hout, nout, cudaHostAllocDefault));
    CK(cudaMalloc(&din, nin)); CK(cudaMalloc(&dout, nout));
    cudaStream_t sk, si, so; CK(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const char* names[3] = {"16 B loads", "32 B sector, lane pairs", "32 B sector, one thread"};
    for (int variant = 0; variant < 3; ++variant) {
        for (int with_dma = 0; with_dma < 2; ++with_dma) {
            const int iters = 50;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(a, sk));
                CK(cudaStreamWaitEvent(si, a, 0)); CK(cudaStreamWaitEvent(so, a, 0));
                for (int it = 0; it < iters; ++it) {
                    if (variant == 0) gather16<<<(rows + 255) / 256, 256, 0, sk>>>(host, idx, out, rows);
                    else if (variant == 1) gather32pair<<<(2 * rows + 255) / 256, 256, 0, sk>>>(host, idx, out, rows);
                    else gather32one<<<(rows + 255) / 256, 256, 0, sk>>>(host, idx, out, rows);
                    if (with_dma) {
                        CK(cudaMemcpyAsync(din, hin, nin, cudaMemcpyHostToDevice, si));
                        CK(cudaMemcpyAsync(hout, dout, nout, cudaMemcpyDeviceToHost, so));
                    }
                }
                CK(cudaEventRecord(b, si)); CK(cudaStreamWaitEvent(sk, b, 0));
                CK(cudaEventRecord(b, so)); CK(cudaStreamWaitEvent(sk, b, 0));
                CK(cudaEventRecord(b, sk));
                CK(cudaEventSynchronize(b));
                float ms; CK(cudaEventElapsedTime(&ms, a, b));
                if (ms < best) best = ms;
            }
            printf("%-26s %s: %.1f us per step (%d rows)\n", names[variant], with_dma ? "+ 2.3 MB H2D + 2.9 MB D2H" : "alone                    ", best * 1e3 / iters, rows);
        }
    }
    {   // the DMA copies alone
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a, si)); CK(cudaStreamWaitEvent(so, a, 0));
            for (int it = 0; it < 50; ++it) {
                CK(cudaMemcpyAsync(din, hin, nin, cudaMemcpyHostToDevice, si));
                CK(cudaMemcpyAsync(hout, dout, nout, cudaMemcpyDeviceToHost, so));
            }
            CK(cudaEventRecord(b, so)); CK(cudaStreamWaitEvent(si, b, 0)); CK(cudaEventRecord(b, si));
            CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
        }
        printf("the DMA copies alone: %.1f us per step\n", best * 1e3 / 50);
    }
    return 0;
}
