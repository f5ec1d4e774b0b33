// The reference's README example (README.md:47-59) and a few more calls, through the C++ mirror.
//   g++ -std=c++17 cpp/example.cpp -o /tmp/lb_example -Llambda_blas_b200 -llbk -Wl,-rpath,$PWD/lambda_blas_b200
#include <cmath>
#include <cstdio>

#include "lambda_blas.hpp"

using namespace lambda_blas;

int main()
{
    try {
        Context ctx(0);
        DVector u(ctx, {1, 2, 3}), v(ctx, {4, 5, 6});
        const float d = sdot(3, u, 1, v, 1);                     // README.md:57-59 prints 32.0
        std::printf("sdot = %g\n", d);
        DVector w = saxpy(3, 2.0f, u, 1, v, 1);                   // pure: v is untouched
        auto hw = w.to_host(), hv = v.to_host();
        std::printf("saxpy = [%g %g %g], v still [%g %g %g]\n", hw[0], hw[1], hw[2], hv[0], hv[1], hv[2]);
        auto sw = sswap(2, u, -1, v, 2);
        auto a = sw.first.to_host(), b = sw.second.to_host();
        std::printf("sswap(2,u,-1,v,2) = [%g %g %g] [%g %g %g]\n", a[0], a[1], a[2], b[0], b[1], b[2]);
        std::printf("snrm2 = %g isamax = %lld sasum = %g\n", snrm2(3, v, 1), (long long)isamax(3, v, 1), sasum(3, v, 1));
        ModGivensRot m = srotmg(1.5f, 0.75f, 2.0f, 0.5f);
        auto r = srotm(m, 3, u, 1, v, 1);
        auto rx = r.first.to_host();
        std::printf("srotm flag %d -> x' = [%g %g %g]\n", m.flag, rx[0], rx[1], rx[2]);
        bool ok = d == 32.0f && hw[0] == 6 && hw[1] == 9 && hw[2] == 12 && hv[0] == 4 && isamax(3, v, 1) == 2 &&
                  a[0] == 6 && a[1] == 4 && a[2] == 3 && b[0] == 2 && b[1] == 5 && b[2] == 1;   // x[1-i] <-> y[2i], Util.hs:118-119
        try { sdot(5, u, 1, v, 1); ok = false; } catch (const Error &e) { ok = ok && e.status == LBK_ERR_RANGE; }
        // the Double instance (ddot, daxpy, idamax: Single.hs:86,224,443) through the same templates
        DVectorD ud(ctx, std::vector<double>{1, 2, 3}), vd(ctx, std::vector<double>{4, 5, 6});
        auto wd = daxpy(3, 2.0, ud, 1, vd, 1).to_host();
        std::printf("ddot = %g daxpy = [%g %g %g] idamax = %lld dnrm2 = %.17g\n", ddot(3, ud, 1, vd, 1), wd[0], wd[1], wd[2],
                    (long long)idamax(3, vd, 1), dnrm2(3, vd, 1));
        ok = ok && ddot(3, ud, 1, vd, 1) == 32.0 && wd[0] == 6 && wd[2] == 12 && idamax(3, vd, 1) == 2 &&
             std::fabs(dnrm2(3, vd, 1) - std::sqrt(77.0)) <= 2e-15 * std::sqrt(77.0);
        {   // ONE host thread, vectors sharded over ranks (lbk_init_multi); device 0 listed twice = two ranks on one GPU
            Context m(std::vector<int>{0, 0});
            std::vector<float> hx(100003), hy(100003);
            for (size_t i = 0; i < hx.size(); ++i) { hx[i] = (float)((i * 37) % 101) - 50.f; hy[i] = (float)((i * 11) % 13) - 6.f; }
            hx[50001] = -77.f; hx[50002] = 77.f;                         // the maximum and its tie on either side of the shard edge
            DVector xs = DVector::sharded(m, hx), ys = DVector::sharded(m, hy);
            double want = 0;
            for (size_t i = 0; i < hx.size(); ++i) want += (double)hx[i] * hy[i];
            DVector zs = saxpy((int64_t)hx.size(), 2.0f, xs, 1, ys, 1);    // pure, sharded like its arguments
            auto hz = zs.to_host();                                      // gathered into one host buffer
            bool mok = xs.is_sharded() && zs.is_sharded() && sdot((int64_t)hx.size(), xs, 1, ys, 1) == (float)want &&
                       isamax((int64_t)hx.size(), xs, 1) == 50001 && hz[50002] == 2.0f * 77.f + hy[50002] && hz.back() == 2.0f * hx.back() + hy.back();
            std::printf("multi-device context: sdot = %g (want %g), isamax = %lld, %s\n", sdot((int64_t)hx.size(), xs, 1, ys, 1), want,
                        (long long)isamax((int64_t)hx.size(), xs, 1), mok ? "ok" : "MISMATCH");
            ok = ok && mok;
        }
        std::printf(ok ? "OK\n" : "MISMATCH\n");
        return ok ? 0 : 1;
    } catch (const Error &e) {
        std::fprintf(stderr, "lbk error %d: %s\n", e.status, e.what());
        return 2;
    }
}
