// Device check of the exact-by-construction arithmetic forms of the window sampler (development tool and GPU test helper):
//   * r_div_nocheck(a, b) (mpp_device.cuh) against the IEEE quotient a / b on random operands in the ranges of the sampler
//     (half extents: 2 * size / (1 + ratio), size in [0, 32], ratio in [0, 1]);
//   * the FMA-corrected reciprocal product of shape_terms (mpp_sweep2.cuh) against __fdiv_rn(s, 3).
// Build and run:  nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_divcheck tools/div_check.cu && tools/_divcheck
#include <cstdio>
#include <cstdlib>
#include "../mpp_cnn_rs_object_detection_b200/csrc/mpp_device.cuh"

__global__ void k_check(const float *a, const float *b, int n, int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (r_div_nocheck(a[i], b[i]) != a[i] / b[i]) atomicAdd(bad, 1);
    const float s = a[i] - 32.0f;  // sums of three mark energies: a few units around zero
    const float q3 = __fmul_rn(s, 0.333333343267440796f);
    if (__fmaf_rn(__fmaf_rn(-3.0f, q3, s), 0.333333343267440796f, q3) != __fdiv_rn(s, 3.0f)) atomicAdd(bad + 1, 1);
}

int main() {
    const int n = 1 << 24;
    float *ha = (float *)malloc(4 * (size_t)n), *hb = (float *)malloc(4 * (size_t)n);
    srand(1);
    for (int i = 0; i < n; ++i) { ha[i] = 64.0f * (float)rand() / (float)RAND_MAX; hb[i] = 1.0f + (float)rand() / (float)RAND_MAX; }
    float *a, *b;
    int *bad, hbad[2] = {0, 0};
    cudaMalloc(&a, 4 * (size_t)n); cudaMalloc(&b, 4 * (size_t)n); cudaMalloc(&bad, 8);
    cudaMemcpy(a, ha, 4 * (size_t)n, cudaMemcpyHostToDevice); cudaMemcpy(b, hb, 4 * (size_t)n, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 8);
    k_check<<<n / 256, 256>>>(a, b, n, bad);
    cudaMemcpy(hbad, bad, 8, cudaMemcpyDeviceToHost);
    const cudaError_t e = cudaGetLastError();
    printf("r_div_nocheck vs IEEE division: %d mismatches of %d; /3 by corrected reciprocal vs __fdiv_rn: %d mismatches of %d (%s)\n", hbad[0], n, hbad[1], n,
           cudaGetErrorString(e));
    return (e == cudaSuccess && hbad[0] == 0 && hbad[1] == 0) ? 0 : 1;
}
