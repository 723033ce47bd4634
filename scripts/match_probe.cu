// Microbenchmark (GPU box): cycles per __match_any_sync as a function of the number of distinct keys in the warp,
// against __ballot_sync / __shfl_sync / __reduce_or_sync.  One warp, dependent chain (latency) and 4 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int distinct, long long* out, unsigned* sink) {
  const int lane = threadIdx.x;
  int key = lane % distinct;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < 1024; i++) {
    unsigned g = __match_any_sync(0xffffffffu, key);
    acc += g;
    key += (g & 1u) ? 0 : 0; /* dependent on the result, value unchanged */
  }
  long long t1 = clock64();
  unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int i = 0; i < 256; i++) {
    a0 += __match_any_sync(0xffffffffu, key + i);
    a1 += __match_any_sync(0xffffffffu, key ^ 1);
    a2 += __match_any_sync(0xffffffffu, key + 2 * i);
    a3 += __match_any_sync(0xffffffffu, key ^ 3);
  }
  long long t2 = clock64();
  unsigned b = 0;
  int x = key;
  for (int i = 0; i < 1024; i++) {
    unsigned m = __ballot_sync(0xffffffffu, x & 1);
    b += m;
    x += (m >> 31);
  }
  long long t3 = clock64();
  for (int i = 0; i < 1024; i++) {
    x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31) + 1;
  }
  long long t4 = clock64();
  if (lane == 0) {
    out[0] = t1 - t0;
    out[1] = t2 - t1;
    out[2] = t3 - t2;
    out[3] = t4 - t3;
  }
  sink[lane] = acc + a0 + a1 + a2 + a3 + b + x;
}
int main() {
  long long* out;
  unsigned* sink;
  cudaMallocManaged(&out, 4 * sizeof(long long));
  cudaMallocManaged(&sink, 32 * sizeof(unsigned));
  for (int d : {1, 2, 4, 8, 16, 32}) {
    probe<<<1, 32>>>(d, out, sink);
    cudaDeviceSynchronize();
    probe<<<1, 32>>>(d, out, sink);
    cudaDeviceSynchronize();
    printf("distinct %2d: match.any dependent %.1f cycles, 4 independent %.1f cycles each, ballot %.1f, shfl %.1f\n", d,
           out[0] / 1024.0, out[1] / 1024.0, out[2] / 1024.0, out[3] / 1024.0);
  }
  return 0;
}
