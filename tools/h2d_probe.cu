// Probe: pinned-host -> device copy rates for the injected-draw buffer [frames][320] float,
// flat versus strided (only the sample windows the receiver reads).  Analysis tool, not product code.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main(int argc, char **argv)
{
    const long n = argc > 1 ? atol(argv[1]) : 1000000;
    const size_t pitch = 320 * 4;
    float *h, *d;
    CK(cudaHostAlloc(&h, n * pitch, cudaHostAllocDefault));
    CK(cudaMalloc(&d, n * pitch));
    for (long i = 0; i < n * 320; i += 1024) h[i] = 1.f;
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a, s);
        CK(cudaMemcpyAsync(d, h, n * pitch, cudaMemcpyHostToDevice, s));
        cudaEventRecord(b, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, a, b);
        printf("flat 1280 B/frame:            %.3f ms  %.1f GB/s\n", ms, n * pitch / ms / 1e6);
        cudaEventRecord(a, s);
        CK(cudaMemcpy2DAsync(d, 1152, h + 32, pitch, 1152, n, cudaMemcpyHostToDevice, s));
        cudaEventRecord(b, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, a, b);
        printf("one strided copy 1152 of 1280: %.3f ms  %.1f GB/s useful, %.2f x flat time\n", ms, n * 1152 / ms / 1e6, 0.0);
        cudaEventRecord(a, s);
        CK(cudaMemcpy2DAsync(d, 1024, h + 32, pitch, 512, n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpy2DAsync(d + 128, 1024, h + 176, pitch, 256, n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpy2DAsync(d + 192, 1024, h + 256, pitch, 256, n, cudaMemcpyHostToDevice, s));
        cudaEventRecord(b, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, a, b);
        printf("three strided copies 1024 of 1280: %.3f ms  %.1f GB/s useful\n", ms, n * 1024 / ms / 1e6);
        {   // the three window copies on three streams at once
            cudaStream_t s3[3]; for (int i = 0; i < 3; ++i) cudaStreamCreate(&s3[i]);
            cudaEvent_t e3[3]; for (int i = 0; i < 3; ++i) cudaEventCreate(&e3[i]);
            cudaEventRecord(a, s);
            for (int i = 0; i < 3; ++i) cudaStreamWaitEvent(s3[i], a, 0);
            CK(cudaMemcpy2DAsync(d, 1024, h + 32, pitch, 512, n, cudaMemcpyHostToDevice, s3[0]));
            CK(cudaMemcpy2DAsync(d + 128, 1024, h + 176, pitch, 256, n, cudaMemcpyHostToDevice, s3[1]));
            CK(cudaMemcpy2DAsync(d + 192, 1024, h + 256, pitch, 256, n, cudaMemcpyHostToDevice, s3[2]));
            for (int i = 0; i < 3; ++i) { cudaEventRecord(e3[i], s3[i]); cudaStreamWaitEvent(s, e3[i], 0); }
            cudaEventRecord(b, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, a, b);
            printf("three strided copies on three streams: %.3f ms  %.1f GB/s useful\n", ms, n * 1024 / ms / 1e6);
            // two halves of the frames on two streams, three copies each
            cudaEventRecord(a, s);
            for (int i = 0; i < 2; ++i) cudaStreamWaitEvent(s3[i], a, 0);
            for (int i = 0; i < 2; ++i) {
                const long f0 = i * (n / 2), c = i ? n - n / 2 : n / 2;
                CK(cudaMemcpy2DAsync(d + f0 * 256, 1024, h + f0 * 320 + 32, pitch, 512, c, cudaMemcpyHostToDevice, s3[i]));
                CK(cudaMemcpy2DAsync(d + f0 * 256 + 128, 1024, h + f0 * 320 + 176, pitch, 256, c, cudaMemcpyHostToDevice, s3[i]));
                CK(cudaMemcpy2DAsync(d + f0 * 256 + 192, 1024, h + f0 * 320 + 256, pitch, 256, c, cudaMemcpyHostToDevice, s3[i]));
            }
            for (int i = 0; i < 2; ++i) { cudaEventRecord(e3[i], s3[i]); cudaStreamWaitEvent(s, e3[i], 0); }
            cudaEventRecord(b, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, a, b);
            printf("two halves on two streams:             %.3f ms  %.1f GB/s useful\n", ms, n * 1024 / ms / 1e6);
            // flat copy split over two streams
            cudaEventRecord(a, s);
            for (int i = 0; i < 2; ++i) cudaStreamWaitEvent(s3[i], a, 0);
            for (int i = 0; i < 2; ++i) CK(cudaMemcpyAsync(d + (size_t)i * (n / 2) * 320, h + (size_t)i * (n / 2) * 320, (n / 2) * pitch, cudaMemcpyHostToDevice, s3[i]));
            for (int i = 0; i < 2; ++i) { cudaEventRecord(e3[i], s3[i]); cudaStreamWaitEvent(s, e3[i], 0); }
            cudaEventRecord(b, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, a, b);
            printf("flat copy split over two streams:      %.3f ms  %.1f GB/s\n", ms, n * pitch / ms / 1e6);
            for (int i = 0; i < 3; ++i) { cudaStreamDestroy(s3[i]); cudaEventDestroy(e3[i]); }
        }
        // chunked like the pipelined sweep
        const long chunk = 131072;
        cudaEventRecord(a, s);
        for (long f = 0; f < n; f += chunk) {
            long c = n - f < chunk ? n - f : chunk;
            CK(cudaMemcpy2DAsync(d + f * 288, 1152, h + f * 320 + 32, pitch, 1152, c, cudaMemcpyHostToDevice, s));
        }
        cudaEventRecord(b, s); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, a, b);
        printf("chunked strided 1152:          %.3f ms  %.1f GB/s useful\n", ms, n * 1152 / ms / 1e6);
    }
    return 0;
}
