// Per-call latency of the C ABI from a compiled host (no interpreter in the way), at the sizes of the
// reference's criterion suite (test/Bench.hs:113: n up to 30 000 on 32 767-element vectors) and at
// BASELINE.json configs[0] (sdot, n = 1e6).  Prints CSV: routine,form,n,mean_us,min_us.
//   g++ -O2 -std=c++17 cpp/latency.cpp -o build/lb_latency -Llambda_blas_b200 -llbk -Wl,-rpath,$PWD/lambda_blas_b200
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <functional>
#include <vector>

#include "lambda_blas.hpp"

using namespace lambda_blas;
using clk = std::chrono::steady_clock;

static void time_it(const char *routine, const char *form, long long n, const std::function<void()> &f, int reps = 300)
{
    for (int i = 0; i < 20; ++i) f();
    std::vector<double> us((size_t)reps);
    for (int i = 0; i < reps; ++i) {
        auto t0 = clk::now();
        f();
        us[(size_t)i] = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
    }
    double sum = 0;
    for (double v : us) sum += v;
    std::printf("%s,%s,%lld,%.2f,%.2f\n", routine, form, n, sum / reps, *std::min_element(us.begin(), us.end()));
}

int main()
{
    try {
        Context ctx(0);
        const long long N = 1000000;
        std::vector<float> h((size_t)N);
        for (long long i = 0; i < N; ++i) h[(size_t)i] = (float)((i * 2654435761u) % 1000) / 1000.f - 0.5f;
        DVector x(ctx, h), y(ctx, h);
        std::printf("routine,form,n,mean_us,min_us\n");
        for (long long n : {1LL, 100LL, 10000LL, 30000LL, N}) {
            time_it("sdot", "scalar to host (lbk_sdot)", n, [&] { (void)sdot(n, x, 1, y, 1); });
            time_it("snrm2", "scalar to host (lbk_snrm2)", n, [&] { (void)snrm2(n, x, 1); });
            time_it("isamax", "scalar to host (lbk_isamax)", n, [&] { (void)isamax(n, x, 1); });
            time_it("saxpy", "pure: new vector + lbk_sync (lbk_saxpy_oop)", n, [&] { DVector w = saxpy(n, 1.0f, x, 1, y, 1); ctx.sync(); });
            time_it("saxpy", "in place + lbk_sync (lbk_saxpy)", n, [&] { ctx.check(lbk_saxpy(ctx.get(), n, 1.0f, x.get(), 1, y.get(), 1)); ctx.sync(); });
            time_it("saxpy", "in place, enqueue only (lbk_saxpy)", n, [&] { ctx.check(lbk_saxpy(ctx.get(), n, 0.0f, x.get(), 1, y.get(), 1)); });
            ctx.sync();
        }
        // what a pure result costs besides its kernel: one allocation and one release per call
        for (long long n : {1024LL, 1LL << 20, 1LL << 24}) {
            time_it("alloc+free", "lbk_vec_alloc + lbk_vec_free, stream busy (no sync)", n, [&] { DVector t(ctx, n); });
            ctx.sync();
            time_it("alloc+free", "lbk_vec_alloc + lbk_vec_free + lbk_sync", n, [&] { { DVector t(ctx, n); } ctx.sync(); });
        }
        // Chains of DEPENDENT kernels (each call reads what the previous one wrote), enqueue only, one sync at the end:
        // microseconds per call.  pdl 0 = ordinary launches, 1 = the library's default (the next kernel is resident and
        // waits before its first load, so the launch latency is hidden).
        std::printf("chain,pdl,n,us_per_call\n");
        DVector res(ctx, 4);
        for (int pdl : {0, 1, 0, 1}) {
            ctx.check(lbk_set_pdl(ctx.get(), pdl));
            for (long long n : {1LL << 16, 1LL << 18, 1LL << 20, 1LL << 22, 1LL << 24}) {
                const int K = 200;
                DVector bx(ctx, n), by(ctx, n);
                ctx.check(lbk_vec_fill(bx.get(), 0.25)); ctx.check(lbk_vec_fill(by.get(), 0.5));
                auto run = [&](const char *name, const std::function<void()> &f) {
                    for (int i = 0; i < 20; ++i) f();
                    ctx.sync();
                    auto t0 = clk::now();
                    for (int i = 0; i < K; ++i) f();
                    ctx.sync();
                    std::printf("%s,%d,%lld,%.2f\n", name, pdl, n, std::chrono::duration<double, std::micro>(clk::now() - t0).count() / K);
                };
                run("saxpy y+=a*x (in place)", [&] { ctx.check(lbk_saxpy(ctx.get(), n, 1e-3f, bx.get(), 1, by.get(), 1)); });
                run("pure saxpy chain: w = saxpy(a, x, w), a new vector per call", [&] { by = saxpy(n, 1e-3f, bx, 1, by, 1); });
                run("sscal then sdot_async", [&] { ctx.check(lbk_sscal(ctx.get(), n, 1.0f, bx.get(), 1));
                                                  ctx.check(lbk_sdot_async(ctx.get(), n, bx.get(), 1, by.get(), 1, res.get())); });
            }
        }
        ctx.check(lbk_set_pdl(ctx.get(), 1));
        // The same chains recorded ONCE into a CUDA graph (lambda_blas::Graph = lbk_capture_begin / lbk_capture_end) and
        // replayed: one host call per replay, the GPU's node-to-node latency per kernel.
        std::printf("graph_chain,n,kernels,eager_us_per_call,replay_us_per_call,host_us_per_replay\n");
        for (long long n : {1LL << 10, 1LL << 14, 1LL << 16, 1LL << 18, 1LL << 20}) {
            const int K = 200, R = 20;
            DVector bx(ctx, n), by(ctx, n);
            ctx.check(lbk_vec_fill(bx.get(), 0.25)); ctx.check(lbk_vec_fill(by.get(), 0.5));
            auto both = [&](const char *name, int per, const std::function<void()> &f) {
                for (int i = 0; i < 20; ++i) f();
                ctx.sync();
                auto t0 = clk::now();
                for (int i = 0; i < K; ++i) f();
                ctx.sync();
                const double eager = std::chrono::duration<double, std::micro>(clk::now() - t0).count() / (K * per);
                Graph g(ctx, [&] { for (int i = 0; i < K; ++i) f(); });
                for (int i = 0; i < 3; ++i) g.launch();
                ctx.sync();
                t0 = clk::now();
                for (int i = 0; i < R; ++i) g.launch();
                const double host = std::chrono::duration<double, std::micro>(clk::now() - t0).count() / R;
                ctx.sync();
                const double replay = std::chrono::duration<double, std::micro>(clk::now() - t0).count() / ((double)R * (double)g.kernels());
                std::printf("%s,%lld,%lld,%.2f,%.2f,%.1f\n", name, n, (long long)g.kernels(), eager, replay, host);
            };
            both("saxpy y+=a*x (in place; dependent)", 1, [&] { ctx.check(lbk_saxpy(ctx.get(), n, 1e-3f, bx.get(), 1, by.get(), 1)); });
            both("sscal then sdot_async", 2, [&] { ctx.check(lbk_sscal(ctx.get(), n, 1.0f, bx.get(), 1));
                                                   ctx.check(lbk_sdot_async(ctx.get(), n, bx.get(), 1, by.get(), 1, res.get())); });
        }
        return 0;
    } catch (const Error &e) {
        std::fprintf(stderr, "lbk error %d: %s\n", e.status, e.what());
        return 2;
    }
}
