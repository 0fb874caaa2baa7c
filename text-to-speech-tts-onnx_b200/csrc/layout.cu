#include "layout.cuh"

namespace b200tts {

namespace {
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const float* ib = in + (long)b * R * C;
  float* ob = out + (long)b * R * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = ib[(long)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) ob[(long)c * R + r] = tile[threadIdx.x][i];
  }
}
}  // namespace

void batched_transpose(const float* in, float* out, int B, int R, int C, cudaStream_t s) {
  if (B <= 0 || R <= 0 || C <= 0) return;
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32), B), block(32, 8);
  B2_CHECK(grid.y <= 65535 && grid.z <= 65535, "batched_transpose grid too large");
  transpose_kernel<<<grid, block, 0, s>>>(in, out, R, C);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace b200tts
