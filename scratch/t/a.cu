__global__ void k(float *g, int n) {
    __shared__ float s[1024];
    s[threadIdx.x] = 0.f; __syncthreads();
    atomicAdd(&s[(threadIdx.x * 7) & 1023], g[threadIdx.x]);
    asm volatile("red.shared.add.f32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(&s[threadIdx.x ^ 1])), "f"(g[threadIdx.x + 32]));
    __syncthreads(); g[threadIdx.x] = s[threadIdx.x];
}
