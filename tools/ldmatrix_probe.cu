#include <cstdio>
#include <cstdint>
__global__ void k(uint32_t *out) {
  __shared__ __align__(128) uint8_t sm[16*16*2];
  for (int i = threadIdx.x; i < 512; i += 32) sm[i] = (uint8_t)i;
  __syncwarp();
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm) + 16 * (threadIdx.x & 15);
  uint32_t r0, r1;
  asm volatile("ldmatrix.sync.aligned.m16n16.x1.trans.shared.b8 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
  out[threadIdx.x * 2] = r0; out[threadIdx.x * 2 + 1] = r1;
}
int main() { uint32_t *d; cudaMalloc(&d, 256); k<<<1,32>>>(d); uint32_t h[64]; cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
  for (int t = 0; t < 32; t++) printf("t%02d: %08x %08x\n", t, h[2*t], h[2*t+1]); return 0; }
