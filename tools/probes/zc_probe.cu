// Development probe: SM-initiated reads of pinned host memory (zero copy) against the copy engine, for the gray-frame upload of
// csb_detect_upload_gray.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a zc_probe.cu -o zc_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// segment = 512 B (one warp x 16 B); seg_ids lists the segments to copy
__global__ void k_gather(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ seg_ids, int n_seg) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int s = w; s < n_seg; s += nw) {
        const size_t o = (size_t)seg_ids[s] * 32 + lane;
        dst[o] = src[o];
    }
}
// 4 segments in flight per warp
__global__ void k_gather4(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ seg_ids, int n_seg) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int s = 4 * w; s < n_seg; s += 4 * nw) {
        uint4 v[4]; size_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { o[k] = (size_t)seg_ids[min(s + k, n_seg - 1)] * 32 + lane; v[k] = src[o[k]]; }
#pragma unroll
        for (int k = 0; k < 4; k++) if (s + k < n_seg) dst[o[k]] = v[k];
    }
}

int main() {
    const size_t bytes = 64ull * 1242 * 375;  // one step's gray frames (29.8 MB)
    const int n_seg_all = (int)(bytes / 512);
    unsigned char* h; CK(cudaHostAlloc(&h, bytes + 512, cudaHostAllocDefault));
    memset(h, 7, bytes);
    unsigned char* d; CK(cudaMalloc(&d, bytes + 512));
    int* d_ids; CK(cudaMalloc(&d_ids, 4 * n_seg_all));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    auto time_it = [&](const char* name, size_t moved, auto fn) {
        for (int i = 0; i < 3; i++) fn();
        CK(cudaStreamSynchronize(st));
        float best = 1e9, sum = 0;
        for (int i = 0; i < 10; i++) {
            CK(cudaEventRecord(a, st)); fn(); CK(cudaEventRecord(b, st)); CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best; sum += ms;
        }
        printf("%-46s best %.3f ms mean %.3f ms  %.1f GB/s (best)\n", name, best, sum / 10, moved / best / 1e6);
    };
    time_it("cudaMemcpyAsync whole batch", bytes, [&] { CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st)); });
    for (int pct : {100, 50, 44, 25}) {
        // runs of 16 segments (8 KB), keep `pct` percent of the runs
        std::vector<int> ids;
        srand(1);
        for (int r = 0; r < n_seg_all / 16; r++) if (rand() % 100 < pct) for (int k = 0; k < 16; k++) ids.push_back(r * 16 + k);
        CK(cudaMemcpy(d_ids, ids.data(), 4 * ids.size(), cudaMemcpyHostToDevice));
        const int n = (int)ids.size();
        for (int ctas : {148, 296, 592, 1184}) {
            char nm[128];
            snprintf(nm, sizeof nm, "zero-copy gather %3d%% (%d segs) %d CTAs x256", pct, n, ctas);
            time_it(nm, (size_t)n * 512, [&] { k_gather<<<ctas, 256, 0, st>>>((const uint4*)h, (uint4*)d, d_ids, n); });
            snprintf(nm, sizeof nm, "zero-copy gather4 %3d%% (%d segs) %d CTAs x256", pct, n, ctas);
            time_it(nm, (size_t)n * 512, [&] { k_gather4<<<ctas, 256, 0, st>>>((const uint4*)h, (uint4*)d, d_ids, n); });
        }
    }
    CK(cudaGetLastError());
    return 0;
}
