// How many kernel launches per second can T host threads push into one GPU (one stream each)?
// nvcc -O2 -arch=sm_100a -o launch_rate launch_rate.cu -lpthread
#include <cstdio>
#include <thread>
#include <vector>
#include <chrono>
#include <cuda_runtime.h>
__global__ void k_empty(int *p) { if (p && threadIdx.x == 1234) *p = 1; }
int main()
{
    cudaFree(0);
    for (int T : {1, 2, 4, 8, 16}) {
        for (int sync_every : {0, 20}) {
            const int N = 20000;
            std::vector<cudaStream_t> st(T);
            for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++) th.emplace_back([&, t] {
                for (int i = 0; i < N; i++) {
                    k_empty<<<1, 32, 0, st[t]>>>(nullptr);
                    if (sync_every && i % sync_every == sync_every - 1) cudaStreamSynchronize(st[t]);
                }
                cudaStreamSynchronize(st[t]);
            });
            for (auto &x : th) x.join();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("threads %2d sync_every %2d: %.0f launches/s aggregate (%.2f us per launch per thread)\n", T, sync_every, T * N / s, s / N * 1e6);
            for (auto &s2 : st) cudaStreamDestroy(s2);
        }
    }
    return 0;
}
