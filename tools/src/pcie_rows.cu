// pcie_rows.cu -- what SM-issued accesses to page-locked HOST memory cost on this box: random rows of 16..128 B
// read (gather into device memory) and written (scatter from device memory), all requests independent.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/pcie_rows tools/src/pcie_rows.cu && gpurun_out/pcie_rows
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// rows of W float4 each; thread t moves float4 (t % W) of row idx[t / W]
template <int W>
__global__ void gather_rows(const float4* __restrict__ host, const int* __restrict__ idx, float4* __restrict__ out, int rows) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * W) return;
    const int r = t / W, q = t % W;
    out[t] = host[(long long)idx[r] * W + q];
}
template <int W>
__global__ void scatter_rows(float4* __restrict__ host, const int* __restrict__ idx, const float4* __restrict__ in, int rows) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * W) return;
    const int r = t / W, q = t % W;
    host[(long long)idx[r] * W + q] = in[t];
}

template <int W>
static void run(int rows, int table_rows, int threads) {
    float4* host; CK(cudaHostAlloc(&host, (size_t)table_rows * W * 16, cudaHostAllocDefault));
    memset(host, 1, (size_t)table_rows * W * 16);
    std::vector<int> hidx(rows);
    unsigned s = 12345u;
    for (int i = 0; i < rows; ++i) { s = s * 1664525u + 1013904223u; hidx[i] = (int)((s >> 8) % (unsigned)table_rows); }
    int* idx; float4* buf;
    CK(cudaMalloc(&idx, rows * 4)); CK(cudaMalloc(&buf, (size_t)rows * W * 16));
    CK(cudaMemcpy(idx, hidx.data(), rows * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(buf, 0, (size_t)rows * W * 16));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = (rows * W + threads - 1) / threads;
    float ms_g = 1e30f, ms_s = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        float ms;
        CK(cudaEventRecord(a)); gather_rows<W><<<grid, threads>>>(host, idx, buf, rows); CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); if (ms < ms_g) ms_g = ms;
        CK(cudaEventRecord(a)); scatter_rows<W><<<grid, threads>>>(host, idx, buf, rows); CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); if (ms < ms_s) ms_s = ms;
    }
    CK(cudaGetLastError());
    printf("rows %7d x %3d B (table %d rows): gather %7.1f us = %6.1f M rows/s %5.1f GB/s | scatter %7.1f us = %6.1f M rows/s %5.1f GB/s\n",
           rows, W * 16, table_rows, ms_g * 1e3, rows / (ms_g * 1e3), rows * W * 16.0 / (ms_g * 1e6),
           ms_s * 1e3, rows / (ms_s * 1e3), rows * W * 16.0 / (ms_s * 1e6));
    CK(cudaFree(idx)); CK(cudaFree(buf)); CK(cudaFreeHost(host));
}

int main() {
    // C2: 64 images x ~1000 candidate rows of 16 B out of 64 x 8649; compact targets: 64 x 128 rows x 2
    for (int rows : {8192, 16384, 65536, 131072}) run<1>(rows, 64 * 8649, 256);
    for (int rows : {65536}) { run<2>(rows, 32 * 8649, 256); run<4>(rows, 16 * 8649, 256); run<8>(rows, 8 * 8649, 256); }
    // sequential rows (the dense tensor read by a kernel instead of the copy engine)
    {
        const int rows = 64 * 8649;
        float4* host; CK(cudaHostAlloc(&host, (size_t)rows * 16, cudaHostAllocDefault)); memset(host, 1, (size_t)rows * 16);
        std::vector<int> hidx(rows); for (int i = 0; i < rows; ++i) hidx[i] = i;
        int* idx; float4* buf; CK(cudaMalloc(&idx, rows * 4)); CK(cudaMalloc(&buf, (size_t)rows * 16));
        CK(cudaMemcpy(idx, hidx.data(), rows * 4, cudaMemcpyHostToDevice));
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        float best = 1e30f, best_s = 1e30f, best_c = 1e30f, ms;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(a)); gather_rows<1><<<(rows + 255) / 256, 256>>>(host, idx, buf, rows); CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
            CK(cudaEventRecord(a)); scatter_rows<1><<<(rows + 255) / 256, 256>>>(host, idx, buf, rows); CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best_s) best_s = ms;
            CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(buf, host, (size_t)rows * 16, cudaMemcpyHostToDevice)); CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best_c) best_c = ms;
        }
        printf("sequential %d x 16 B: kernel read %.1f us %.1f GB/s | kernel write %.1f us %.1f GB/s | copy engine H2D %.1f us %.1f GB/s\n", rows,
               best * 1e3, rows * 16.0 / (best * 1e6), best_s * 1e3, rows * 16.0 / (best_s * 1e6), best_c * 1e3, rows * 16.0 / (best_c * 1e6));
    }
    return 0;
}
