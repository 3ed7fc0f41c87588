// tools/l2bw.cu -- L2 read-bandwidth microbenchmark ("B200 L2 bandwidth: not
// yet measured -- builder must microbenchmark").  Every thread block streams the whole
// working set (LDG.128, .cg = L2 only) several times; working sets below the L2 capacity
// measure L2 bandwidth, larger ones fall to HBM.   nvcc -O3 -arch=sm_100a -cudart shared l2bw.cu -o /tmp/l2bw   (build outside the repo tree)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024) rd(const float4* __restrict__ p, size_t n, int reps, float* out)
{
        float acc = 0.f;
        // every block walks the WHOLE working set, each starting at its own offset so that
        // different blocks touch different lines at any instant (no request merging)
        const size_t start = (n / gridDim.x) * blockIdx.x;
        for (int r = 0; r < reps; ++r)
                for (size_t k = threadIdx.x; k < n; k += blockDim.x) {
                        size_t i = start + k;
                        if (i >= n)
                                i -= n;
                        float4 v = __ldcg(p + i);
                        acc += v.x + v.y + v.z + v.w;
                }
        if (acc == 123.456f)
                *out = acc;
}
int main()
{
        cudaDeviceProp pr;
        cudaGetDeviceProperties(&pr, 0);
        printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"results\": [", pr.name, pr.multiProcessorCount, pr.l2CacheSize);
        float* out;
        cudaMalloc(&out, 4);
        bool first = true;
        for (size_t mb : { 8, 16, 32, 48, 64, 96, 256, 1024 }) {
                size_t bytes = mb << 20, n = bytes / 16;
                float4* p;
                cudaMalloc(&p, bytes);
                cudaMemset(p, 0, bytes);
                int reps = 1;
                cudaEvent_t a, b;
                cudaEventCreate(&a);
                cudaEventCreate(&b);
                float best = 1e30f;
                for (int it = 0; it < 5; ++it) {
                        cudaEventRecord(a);
                        rd<<<pr.multiProcessorCount * 2, 1024>>>(p, n, reps, out);
                        cudaEventRecord(b);
                        cudaEventSynchronize(b);
                        float ms;
                        cudaEventElapsedTime(&ms, a, b);
                        if (ms < best) best = ms;
                }
                printf("%s{\"working_set_mb\": %zu, \"gb_per_s\": %.1f}", first ? "" : ", ", mb, (double)bytes * reps * pr.multiProcessorCount * 2 / best / 1e6);
                first = false;
                cudaFree(p);
        }
        printf("]}\n");
        return 0;
}
