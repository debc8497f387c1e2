// pcie_pattern.cu -- what the PCIe link of this box does with the step's copy pattern (no kernels).
//   nvcc -O2 -o gpurun_out/pcie_pattern tools/src/pcie_pattern.cu && gpurun_out/pcie_pattern
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

static double run(const std::vector<size_t>& h2d, const std::vector<size_t>& d2h, int steps, unsigned flags_in, bool one_alloc) {
    cudaStream_t si, so;
    CK(cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking));
    size_t tin = 0, tout = 0;
    for (size_t b : h2d) tin += b;
    for (size_t b : d2h) tout += b;
    char *hin = nullptr, *hout = nullptr, *din, *dout;
    std::vector<char*> hins;
    if (tin) {
        if (one_alloc) { CK(cudaHostAlloc(&hin, tin, flags_in)); memset(hin, 1, tin); }
        else for (size_t b : h2d) { char* p; CK(cudaHostAlloc(&p, b, flags_in)); memset(p, 1, b); hins.push_back(p); }
    }
    if (tout) CK(cudaHostAlloc(&hout, tout, cudaHostAllocDefault));
    CK(cudaMalloc(&din, tin + 256)); CK(cudaMalloc(&dout, tout + 256));
    cudaEvent_t a, b, c;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreate(&c));
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, si));
        CK(cudaStreamWaitEvent(so, a, 0));
        for (int s = 0; s < steps; ++s) {
            size_t o = 0; int k = 0;
            for (size_t by : h2d) { CK(cudaMemcpyAsync(din + o, one_alloc ? hin + o : hins[k], by, cudaMemcpyHostToDevice, si)); o += by; ++k; }
            o = 0;
            for (size_t by : d2h) { CK(cudaMemcpyAsync(hout + o, dout + o, by, cudaMemcpyDeviceToHost, so)); o += by; }
        }
        CK(cudaEventRecord(b, so));
        CK(cudaStreamWaitEvent(si, b, 0));
        CK(cudaEventRecord(c, si));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, c));
        if (ms < best) best = ms;
    }
    if (hin) cudaFreeHost(hin);
    for (char* p : hins) cudaFreeHost(p);
    if (hout) cudaFreeHost(hout);
    cudaFree(din); cudaFree(dout);
    cudaStreamDestroy(si); cudaStreamDestroy(so);
    return best * 1e3 / steps;
}

int main() {
    const size_t REG = 64ull * 8649 * 16, CLS = 64ull * 8649 * 4, GT = 64 * 50 * 16, GL = 64 * 50 * 4, OUT = 64 * 300 * 24 + 256;
    struct Case { const char* name; std::vector<size_t> in, out; };
    std::vector<Case> cases = {
        {"h2d only: gt gl reg cls", {GT, GL, REG, CLS}, {}},
        {"h2d only: one block", {GT + GL + REG + CLS}, {}},
        {"d2h only: deltas labels outs", {}, {REG, CLS, OUT}},
        {"duplex: 4 in / 3 out (step pattern)", {GT, GL, REG, CLS}, {REG, CLS, OUT}},
        {"duplex: 1 in / 1 out", {GT + GL + REG + CLS}, {REG + CLS + OUT}},
        {"duplex: 2 in / 2 out", {REG, CLS}, {REG, CLS}},
        {"duplex: 8 in / 8 out (4 chunks)", {REG / 4, CLS / 4, REG / 4, CLS / 4, REG / 4, CLS / 4, REG / 4, CLS / 4},
                                             {REG / 4, CLS / 4, REG / 4, CLS / 4, REG / 4, CLS / 4, REG / 4, CLS / 4}},
        {"h2d only 2.8MB", {2800000}, {}},
        {"h2d only 1.4MB", {1400000}, {}},
        {"h2d only 5.5MB", {5500000}, {}},
    };
    for (auto& cs : cases) {
        double t0 = run(cs.in, cs.out, 40, cudaHostAllocDefault, true);
        double t1 = run(cs.in, cs.out, 40, cudaHostAllocWriteCombined, true);
        double t2 = run(cs.in, cs.out, 40, cudaHostAllocDefault, false);
        printf("%-40s default %7.1f us/step   write-combined-in %7.1f   separate allocs %7.1f\n", cs.name, t0, t1, t2);
    }
    return 0;
}
