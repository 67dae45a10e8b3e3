// dev tool: host memory bandwidth of the float32 -> float64 widening pass (threads, AVX2 streaming stores)
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static void widen(const float *src, double *dst, size_t n, bool stream)
{
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m256 v = _mm256_loadu_ps(src + i);
        if (stream) {
            _mm256_stream_pd(dst + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
            _mm256_stream_pd(dst + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
        } else {
            _mm256_storeu_pd(dst + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
            _mm256_storeu_pd(dst + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
        }
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
int main(int argc, char **argv)
{
    const size_t n = 200u * 1000 * 1000;
    float *src = (float *)aligned_alloc(64, n * 4);
    double *dst = (double *)aligned_alloc(64, n * 8);
    memset(src, 1, n * 4);
    memset(dst, 0, n * 8);
    for (int threads : {1, 2, 4, 8, 16, 32}) {
        if (threads > (int)std::thread::hardware_concurrency() * 2) break;
        for (int stream = 0; stream < 2; ++stream) {
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> ts;
                const size_t per = n / threads;
                for (int t = 0; t < threads; ++t) ts.emplace_back(widen, src + t * per, dst + t * per, per, stream != 0);
                for (auto &t : ts) t.join();
                best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            printf("threads %2d stream %d: %.2f ms  (%.1f GB/s written, %.1f GB/s total)\n", threads, stream, best * 1e3, n * 8 / best / 1e9, n * 12 / best / 1e9);
        }
    }
    // plain memcpy of 1.6 GB
    {
        char *a = (char *)dst; 
        char *b = (char *)aligned_alloc(64, n * 8);
        memset(b, 0, n * 8);
        for (int threads : {1, 8, 16}) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> ts;
            const size_t per = n * 8 / threads;
            for (int t = 0; t < threads; ++t) ts.emplace_back([=] { memcpy(b + t * per, a + t * per, per); });
            for (auto &t : ts) t.join();
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("memcpy threads %2d: %.2f ms (%.1f GB/s copied)\n", threads, dt * 1e3, n * 8 / dt / 1e9);
        }
        // first-touch cost of a fresh 1.6 GB allocation
        for (int threads : {1, 16}) {
            char *c = (char *)malloc(n * 8);
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> ts;
            const size_t per = n * 8 / threads;
            for (int t = 0; t < threads; ++t) ts.emplace_back([=] { for (size_t i = 0; i < per; i += 4096) c[t * per + i] = 1; });
            for (auto &t : ts) t.join();
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("first touch of 1.6 GB, threads %2d: %.2f ms\n", threads, dt * 1e3);
            free(c);
        }
    }
    FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    if (f) { char buf[128] = {0}; fgets(buf, 127, f); printf("THP: %s", buf); fclose(f); }
    return 0;
}
