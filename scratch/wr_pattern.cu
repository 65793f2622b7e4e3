// Memory-system ceiling for the featuriser's write pattern: one warp per env writes its slices of the
// four output tensors with 128-bit streaming stores and does nothing else.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(128) wr(float* obs, float* vec, float* gst, float* gvec, int n, int mode) {
  int lane = threadIdx.x & 31, w = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= n) return;
  float4 v = make_float4(1.f, 0.f, 0.f, (float)w);
  float4* o = (float4*)(obs + (size_t)w * 3000);
  for (int i = lane; i < 750; i += 32) __stcs(o + i, v);
  float4* g = (float4*)(gst + (size_t)w * 400);
  for (int i = lane; i < 100; i += 32) __stcs(g + i, v);
  float4* ve = (float4*)(vec + (size_t)w * 260);
  for (int i = lane; i < 65; i += 32) __stcs(ve + i, v);
  if (mode == 0) {  // gvec rows are 1301 floats: misaligned, like the real kernel (scalar head/tail)
    size_t g0 = (size_t)w * 1301; int off = g0 & 3, head = (4 - off) & 3;
    float* gv = gvec + g0;
    if (lane < head) __stcs(gv + lane, 1.f);
    float4* b = (float4*)(gv + head); int body = (1301 - head) >> 2;
    for (int i = lane; i < body; i += 32) __stcs(b + i, v);
    int done = head + body * 4; if (lane < 1301 - done) __stcs(gv + done + lane, 2.f);
  } else {
    float4* b = (float4*)(gvec + (size_t)w * 1304);
    for (int i = lane; i < 326; i += 32) __stcs(b + i, v);
  }
}
int main() {
  int n = 65536; float *obs, *vec, *gst, *gvec;
  cudaMalloc(&obs, (size_t)n * 3000 * 4); cudaMalloc(&vec, (size_t)n * 260 * 4); cudaMalloc(&gst, (size_t)n * 400 * 4); cudaMalloc(&gvec, (size_t)n * 1304 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 2; ++mode) {
    for (int i = 0; i < 5; ++i) wr<<<n / 4, 128>>>(obs, vec, gst, gvec, n, mode);
    cudaEventRecord(a);
    for (int i = 0; i < 50; ++i) wr<<<n / 4, 128>>>(obs, vec, gst, gvec, n, mode);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 50;
    double bytes = (double)n * (3000 + 260 + 400 + 1301) * 4;
    printf("mode %d: %.1f us  %.0f GB/s  (%s)\n", mode, ms * 1e3, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
