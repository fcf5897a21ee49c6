// Cost of one cluster-wide exchange step on sm_100a, cycles per iteration (256 threads per CTA, 1 CTA per SM):
//  V0 cg::cluster.sync()                      V1 fence.cta + arrive.relaxed + wait.acquire
//  V2 remote store + cg sync (push)           V3 local store + light barrier + remote load (pull)
//  V5 remote store + light barrier (push, not covered by the memory model: mismatches are counted)
#include <cstdio>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ void light_sync() {
    asm volatile("fence.acq_rel.cta;\n" ::: "memory");
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
template<int V>
__global__ void k(long long *out, int iters, int *bad) {
    cg::cluster_group cl = cg::this_cluster();
    const int cs = cl.num_blocks(), rk = cl.block_rank(), t = threadIdx.x;
    __shared__ double buf[2][8][256];
    double acc = 0;
    int mism = 0;
    cl.sync();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int par = it & 1;
        const double val = it * 1000.0 + rk * 10 + 1;
        if (V == 0) cl.sync();
        else if (V == 1) light_sync();
        else if (V == 2 || V == 5) {
            for (int r = 0; r < cs; ++r) cl.map_shared_rank(&buf[par][rk][t], r)[0] = val;
            if (V == 2) cl.sync(); else light_sync();
            for (int r = 0; r < cs; ++r) { const double v = buf[par][r][t]; acc += v; mism += (v != it * 1000.0 + r * 10 + 1); }
        } else {
            buf[par][0][t] = val;
            light_sync();
            for (int r = 0; r < cs; ++r) { const double v = cl.map_shared_rank(&buf[par][0][t], r)[0]; acc += v; mism += (v != it * 1000.0 + r * 10 + 1); }
        }
    }
    const long long t1 = clock64();
    cl.sync();
    if (t == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
    if (acc == -1.0) out[1] = 1;
    if (mism) atomicAdd(bad, mism);
}
template<int V> void run(int cs) {
    long long *out; int *bad; cudaMalloc(&out, 16); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs * 32); cfg.blockDim = dim3(256);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k<V>, out, 2000, bad);
    long long h[2]; int hb; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("V%d cluster %d: %lld cycles/iter, mismatches %d (%s)\n", V, cs, h[0], hb, cudaGetErrorString(cudaGetLastError()));
}
int main() { for (int cs : {2, 4}) { run<0>(cs); run<1>(cs); run<2>(cs); run<3>(cs); run<5>(cs); } return 0; }
