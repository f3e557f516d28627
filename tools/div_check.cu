#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ float r_div_nocheck(float a, float b) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(b));
    y = __fmaf_rn(y, __fmaf_rn(-b, y, 1.0f), y);
    const float q = __fmul_rn(a, y);
    return __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
}
__global__ void k(const float *a, const float *b, int n, int *bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { if (r_div_nocheck(a[i], b[i]) != a[i] / b[i]) atomicAdd(bad, 1); }
}
int main() {
    const int n = 1 << 24;
    float *ha = (float *)malloc(4 * n), *hb = (float *)malloc(4 * n);
    srand(1);
    for (int i = 0; i < n; ++i) { ha[i] = 64.0f * rand() / RAND_MAX; hb[i] = 1.0f + (float)rand() / RAND_MAX; }
    float *a, *b; int *bad, hbad = 0;
    cudaMalloc(&a, 4 * n); cudaMalloc(&b, 4 * n); cudaMalloc(&bad, 4);
    cudaMemcpy(a, ha, 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(b, hb, 4 * n, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 4);
    k<<<n / 256, 256>>>(a, b, n, bad);
    cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    printf("mismatches vs IEEE division: %d of %d (%s)\n", hbad, n, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
