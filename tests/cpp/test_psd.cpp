// C++ twin of the reference's unit tests for the hot path, written against include/sspsd.hpp so that it
// reads like the Rust original (reference src/psd.rs:599-644 `test`, src/psd.rs:562-597 `exact`,
// src/var.rs:52-60 `basic`).  Exit code 0 = all assertions hold, 3 = no CUDA device (no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sspsd.hpp"

#define REQUIRE(c)                                                        \
    do {                                                                  \
        if (!(c)) {                                                       \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                     \
        }                                                                 \
    } while (0)

int main()
{
    using namespace sspsd;
    constexpr size_t N = 1 << 9;
    try {
        // make uniform noise with zero mean and rms = 1 (psd.rs:603-612), seeded LCG instead of rand::random
        std::vector<float> x(1 << 16);
        uint64_t s = 0x7654321;
        double xm = 0, xv = 0;
        for (auto& v : x) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            v = ((float)((s >> 40) & 0xffffff) / 16777216.0f - 0.5f) * std::sqrt(12.0f);
            xm += v;
            xv += (double)v * v;
        }
        REQUIRE(std::fabs(xm / x.size()) < 10.0 / std::sqrt((double)x.size()));
        REQUIRE(std::fabs(xv / x.size() - 1.0) < 10.0 / std::sqrt((double)x.size()));

        Psd<N> st(Window::Hann);
        std::vector<float> y(x.size() >> 3);
        size_t ny = st.process(x.data(), x.size(), y.data(), y.size());
        REQUIRE(ny == (x.size() >> 3) - 115);  // psd.rs:622 with hbf_dec_response_length(3) = 115
        float g = 1.0f / st.gain();
        auto sp = st.spectrum();
        REQUIRE(st.count() == 255);
        for (float p : sp) REQUIRE(std::fabs(p * g * 0.5f - 1.0f) < 10.0f / std::sqrt((float)st.count()));  // psd.rs:626-632

        PsdCascade<N> d;
        d.process(x);
        auto [p, b] = d.psd(MergeOpts());
        REQUIRE(b.size() >= 2);
        for (const auto& bi : b)
            for (size_t k = bi.start; k < bi.start + (bi.bins.second - bi.bins.first) && bi.include; ++k)
                REQUIRE(std::fabs(p[k] * 0.5f - 1.0f) < 10.0f / std::sqrt((float)bi.count));  // psd.rs:637-643
        auto f = Break::frequencies(b);
        REQUIRE(f.size() == p.size() && f.front() == 0.0f && f.back() == 0.5f);  // psd.rs:324-325

        // clone keeps streaming state (psd.rs:399)
        PsdCascade<N> d2(d);
        d.process(x);
        d2.process(x);
        auto pa = d.psd().first, pb = d2.psd().first;
        REQUIRE(pa.size() == pb.size());
        for (size_t k = 0; k < pa.size(); ++k) REQUIRE(std::fabs(pa[k] - pb[k]) <= 1e-5f * pa[k]);

        // the commented-out exact test generalised: Hann, x = ones -> [4N/3, N/3, 0, ...] (psd.rs:562-597)
        Psd<N> e(Window::Hann);
        std::vector<float> ones(N, 1.0f), y2(N);
        e.process(ones.data(), N, y2.data(), y2.size());
        auto se = e.spectrum();
        float ge = e.gain();
        REQUIRE(std::fabs(se[0] / ge - 4.0f * N / 3) < 1e-5f * N && std::fabs(se[1] / ge - N / 3.0f) < 1e-5f * N);

        // Detrend::Linear is unimplemented!() in the reference (psd.rs:110)
        bool threw = false;
        try {
            d.set_detrend(Detrend::Linear);
        } catch (const Error& er) {
            threw = er.status == SSPSD_EUNIMPLEMENTED;
        }
        REQUIRE(threw);

        // device-resident white noise source (source.rs:104-118) through a cascade: flat PSD (psd.rs:637-643)
        {
            PsdCascade<N> dn;
            Source src = Source::noise(0);
            src.feed(dn, 1 << 20);
            src.feed(dn, 12345);
            REQUIRE(src.position() == (1 << 20) + 12345);
            auto [pn, bn] = dn.psd(MergeOpts());
            for (const auto& bi : bn)
                for (size_t k = bi.start; k < bi.start + (bi.bins.second - bi.bins.first) && bi.include; ++k)
                    REQUIRE(std::fabs(pn[k] * 0.5f - 1.0f) < 10.0f / std::sqrt((float)bi.count));
        }

        // the receiver loop over traces (bin/psd.rs:170-183) on a group: three traces on two ranks that share GPU 0
        {
            Group<N> grp({0, 0});
            grp.set_detrend(Detrend::Midpoint);
            std::vector<std::vector<float>> tr(3, x);
            for (size_t c = 0; c < tr.size(); ++c)
                for (auto& v : tr[c]) v *= (float)(c + 1);
            for (size_t c = 0; c < tr.size(); ++c) grp.process((uint32_t)c, tr[c]);
            for (size_t c = 0; c < tr.size(); ++c) {
                PsdCascade<N> one;
                one.set_detrend(Detrend::Midpoint);
                one.process(tr[c]);
                auto [pg, bg] = grp.psd((uint32_t)c);
                auto [po, bo] = one.psd();
                REQUIRE(pg.size() == po.size() && bg.size() == bo.size());
                for (size_t k = 0; k < pg.size(); ++k) REQUIRE(std::fabs(pg[k] - po[k]) <= 1e-5f * po[k]);
            }
            // ONE stream cut into three time chunks: same spectrum as the sequential cascade
            Group<N> tg({0, 0, 0}, Group<N>::Shard::Time);
            std::vector<float> longx;
            for (int r = 0; r < 40; ++r) longx.insert(longx.end(), x.begin(), x.end());
            tg.time_plan(longx.size());
            tg.time_process_all(longx);
            tg.time_finish();
            PsdCascade<N> seq;
            seq.process(longx);
            auto [pt, bt] = tg.psd();
            auto [ps, bs] = seq.psd();
            REQUIRE(pt.size() == ps.size() && bt.size() == bs.size());
            for (size_t i = 0; i < bt.size(); ++i) REQUIRE(bt[i].count == bs[i].count && bt[i].pending == bs[i].pending);
            for (size_t k = 0; k < pt.size(); ++k) REQUIRE(std::fabs(pt[k] - ps[k]) <= 5e-5f * ps[k]);
        }

        // Trace::plot + Trapezoidal (bin/psd.rs:96-157): trapezoids 0.2 + 0.3 + 0.8 between four points
        {
            Trace t{"t", {}, {2.0f, 2.0f, 4.0f, 4.0f}, {0.0f, 0.1f, 0.2f, 0.4f}};
            auto [integral, pts] = t.plot();
            REQUIRE(std::fabs(integral - std::sqrt(1.3f)) < 1e-6f && pts.size() == 3);
        }

        // var.rs:52-60
        float v = Var().eval({1000.0f, 100.0f, 1.2f, 3.4f, 5.6f}, {0.0f, 1.0f, 3.0f, 6.0f, 9.0f}, 2.7f);
        REQUIRE(std::fabs(0.13478442f - v) < 1e-6f);
    } catch (const sspsd::Error& e) {
        std::fprintf(stderr, "sspsd error %d: %s\n", e.status, e.what());
        return e.status == SSPSD_ECUDA ? 3 : 2;
    }
    std::puts("cpp tests ok");
    return 0;
}
