// Probe: aggregate pinned-host -> device copy rate when N GPUs of one box copy at once -- what bounds the END-TO-END
// (host-buffer) sweep at N > 1.  Per configuration: every GPU copies `mb` MiB `reps` times from its own pinned buffer,
// all GPUs started together; reports per-GPU and aggregate GB/s.
//   tools/h2d_multi_probe <n_gpus> [mb=1024] [reps=5]
// Configurations: GPUs one at a time (the per-GPU ceiling) / all together from one process (one thread per GPU) with
// cudaHostAllocDefault, cudaHostAllocPortable, cudaHostAllocWriteCombined, and with the buffer first-touched by the copying
// thread.  (One process per GPU is what bench.py does under torchrun: its e2e line is that measurement.)
// Analysis tool, not product code.
#include <cuda_runtime.h>
#include <pthread.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Job { int dev; size_t bytes; int reps; unsigned flags; bool touch_in_thread; double seconds; float *h; float *d; pthread_barrier_t *bar; };

static void *worker(void *p)
{
    Job *j = (Job *)p;
    cudaSetDevice(j->dev);
    if (cudaHostAlloc((void **)&j->h, j->bytes, j->flags) != cudaSuccess) { printf("cudaHostAlloc failed on %d\n", j->dev); j->seconds = -1; pthread_barrier_wait(j->bar); return nullptr; }
    cudaMalloc((void **)&j->d, j->bytes);
    if (j->touch_in_thread || true) memset(j->h, 1, j->bytes);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaMemcpyAsync(j->d, j->h, j->bytes, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s);     // warm
    pthread_barrier_wait(j->bar);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < j->reps; ++r) cudaMemcpyAsync(j->d, j->h, j->bytes, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    j->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cudaStreamDestroy(s); cudaFree(j->d); cudaFreeHost(j->h);
    return nullptr;
}

static double run(const std::vector<int> &devs, size_t bytes, int reps, unsigned flags, const char *label)
{
    pthread_barrier_t bar; pthread_barrier_init(&bar, nullptr, (unsigned)devs.size());
    std::vector<Job> jobs(devs.size()); std::vector<pthread_t> th(devs.size());
    for (size_t i = 0; i < devs.size(); ++i) { jobs[i] = Job{devs[i], bytes, reps, flags, true, 0, nullptr, nullptr, &bar}; pthread_create(&th[i], nullptr, worker, &jobs[i]); }
    for (auto &t : th) pthread_join(t, nullptr);
    double agg = 0, slow = 0;
    for (auto &j : jobs) { if (j.seconds > slow) slow = j.seconds; }
    for (auto &j : jobs) agg += (double)bytes * reps / j.seconds / 1e9;
    printf("%-46s %zu GPU(s): aggregate %.1f GB/s (sum of per-GPU rates), %.1f GB/s by the slowest GPU's clock; per GPU:", label, devs.size(), agg,
           (double)bytes * reps * devs.size() / slow / 1e9);
    for (auto &j : jobs) printf(" %.1f", (double)bytes * reps / j.seconds / 1e9);
    printf("\n");
    pthread_barrier_destroy(&bar);
    return agg;
}

int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 1, have = 0;
    const size_t bytes = (size_t)(argc > 2 ? atol(argv[2]) : 1024) << 20;
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    cudaGetDeviceCount(&have);
    if (n > have) n = have;
    printf("h2d_multi_probe: %d GPU(s), %zu MiB per copy, %d copies each\n", n, bytes >> 20, reps);
    for (int d = 0; d < n; ++d) { std::vector<int> one{d}; char l[64]; snprintf(l, sizeof l, "GPU %d alone (cudaHostAllocDefault)", d); run(one, bytes, reps, cudaHostAllocDefault, l); }
    std::vector<int> all; for (int d = 0; d < n; ++d) all.push_back(d);
    for (int k = 2; k <= n; k *= 2) { std::vector<int> sub(all.begin(), all.begin() + k); run(sub, bytes, reps, cudaHostAllocDefault, "together, cudaHostAllocDefault"); }
    run(all, bytes, reps, cudaHostAllocPortable, "together, cudaHostAllocPortable");
    run(all, bytes, reps, cudaHostAllocWriteCombined, "together, cudaHostAllocWriteCombined");
    return 0;
}
