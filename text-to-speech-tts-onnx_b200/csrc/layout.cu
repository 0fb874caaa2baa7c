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
__global__ void conv_w_permute_kernel(const float* __restrict__ W, float* __restrict__ out, int ng, int cg, int k, int groups, int n_major) {
  const long total = (long)groups * ng * cg * k;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int n, c;
    long rest;
    if (n_major) { n = (int)(i % ng); c = (int)((i / ng) % cg); rest = i / ((long)ng * cg); }
    else { c = (int)(i % cg); n = (int)((i / cg) % ng); rest = i / ((long)ng * cg); }
    const int j = (int)(rest % k), g = (int)(rest / k);
    out[i] = W[((long)(g * ng + n) * cg + c) * k + j];
  }
}
__global__ void transpose_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C, int ldo) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)C * ldo) return;
  const int c = (int)(i / ldo), r = (int)(i - (long)c * ldo);
  out[i] = r < R ? in[(long)r * C + c] : 0.f;
}
}  // namespace

void conv_weight_permute(const float* W, float* out, int Cout, int cin_g, int k, int groups, int n_major, cudaStream_t s) {
  const long total = (long)Cout * cin_g * k;
  conv_w_permute_kernel<<<ceil_div(total, 256) > 4096 ? 4096 : ceil_div(total, 256), 256, 0, s>>>(W, out, Cout / groups, cin_g, k, groups, n_major);
  B2_LAUNCH_CHECK();
}
void transpose_pad(const float* in, float* out, int R, int C, int ldo, cudaStream_t s) {
  transpose_pad_kernel<<<ceil_div((long)C * ldo, 256), 256, 0, s>>>(in, out, R, C, ldo);
  B2_LAUNCH_CHECK();
}

void batched_transpose(const float* in, float* out, int B, int R, int C, cudaStream_t s) {
  if (B <= 0 || R <= 0 || C <= 0) return;
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32), B), block(32, 8);
  B2_CHECK(grid.y <= 65535 && grid.z <= 65535, "batched_transpose grid too large");
  transpose_kernel<<<grid, block, 0, s>>>(in, out, R, C);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace b200tts
