// Developer aid: does a profiler keep the cluster shape of a cooperative + cluster launch?  (round-1 driver failure)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned* out) {
  unsigned n, r, id;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(n));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(id));
  if (threadIdx.x == 0) { out[blockIdx.x * 3] = n; out[blockIdx.x * 3 + 1] = r; out[blockIdx.x * 3 + 2] = id; }
}
int main() {
  unsigned* d; cudaMalloc(&d, 16 * 3 * 4);
  for (int coop = 0; coop < 2; ++coop) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(16); cfg.blockDim = dim3(32);
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe, d);
    unsigned h[48]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("coop=%d launch=%s: block 9 -> cluster_nctaid %u ctarank %u clusterid %u\n", coop, cudaGetErrorString(e), h[27], h[28], h[29]);
  }
  return 0;
}
