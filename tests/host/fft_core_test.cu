// Host-side check of the __host__ __device__ FFT building blocks (no GPU needed).
// Build: nvcc -std=c++17 -O1 -o /tmp/fft_core_test tests/host/fft_core_test.cu && /tmp/fft_core_test
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../zaf-python_b200/csrc/fft_core.cuh"

namespace zafb {
std::string& last_error_ref() { static std::string s; return s; }
int fail(int code, const char*, ...) { return code; }
std::atomic<int64_t> g_launches{0};
}
using namespace zafb;
typedef std::complex<double> cd;

static std::vector<cd> dft(const std::vector<cd>& x) {
    size_t n = x.size();
    std::vector<cd> y(n);
    for (size_t k = 0; k < n; ++k) {
        cd s = 0;
        for (size_t j = 0; j < n; ++j) s += x[j] * std::polar(1.0, -2.0 * M_PI * double((k * j) % n) / double(n));
        y[k] = s;
    }
    return y;
}

template <int N>
static double check_reg() {
    float2 v[N];
    std::vector<cd> x(N);
    for (int i = 0; i < N; ++i) {
        v[i] = make_float2(float(rand()) / RAND_MAX - 0.5f, float(rand()) / RAND_MAX - 0.5f);
        x[i] = cd(v[i].x, v[i].y);
    }
    auto y = dft(x);
    fft_reg<N>(v);
    double err = 0, scale = 0;
    for (int k = 0; k < N; ++k) {
        float2 g = v[bitrev(k, clog2(N))];
        err = std::max(err, std::abs(cd(g.x, g.y) - y[k]));
        scale = std::max(scale, std::abs(y[k]));
    }
    return err / scale;
}

static double check_stockham(int log2m) {
    int M = 1 << log2m;
    std::vector<float2> a(M), b(M), tw(M);
    std::vector<cd> x(M);
    for (int i = 0; i < M; ++i) {
        a[i] = make_float2(float(rand()) / RAND_MAX - 0.5f, float(rand()) / RAND_MAX - 0.5f);
        x[i] = cd(a[i].x, a[i].y);
        tw[i] = make_float2(float(cos(-2.0 * M_PI * i / M)), float(sin(-2.0 * M_PI * i / M)));
    }
    auto y = dft(x);
    int radix[16];
    int np = stockham_schedule(log2m, radix);
    float2 *pa = a.data(), *pb = b.data();
    int Ns = 1;
    for (int p = 0; p < np; ++p) {
        int R = radix[p];
        for (int j = 0; j < M / R; ++j) {
            if (R == 4) stockham_pass<4>(pa, pb, tw.data(), M, Ns, j);
            else if (R == 8) stockham_pass<8>(pa, pb, tw.data(), M, Ns, j);
            else stockham_pass<2>(pa, pb, tw.data(), M, Ns, j);
        }
        std::swap(pa, pb);
        Ns *= R;
    }
    double err = 0, scale = 0;
    for (int k = 0; k < M; ++k) {
        err = std::max(err, std::abs(cd(pa[k].x, pa[k].y) - y[k]));
        scale = std::max(scale, std::abs(y[k]));
    }
    return err / scale;
}

int main() {
    int bad = 0;
    double e;
    // compile-time trig vs libm
    double terr = 0;
    terr = std::max(terr, std::fabs(double(Tw<1, 32>::re) - cos(2 * M_PI / 32)));
    terr = std::max(terr, std::fabs(double(Tw<7, 32>::im) + sin(2 * M_PI * 7 / 32)));
    terr = std::max(terr, std::fabs(double(Tw<13, 64>::re) - cos(2 * M_PI * 13 / 64)));
    terr = std::max(terr, std::fabs(double(Tw<29, 64>::im) + sin(2 * M_PI * 29 / 64)));
    terr = std::max(terr, std::fabs(double(Tw<45, 64>::re) - cos(2 * M_PI * 45 / 64)));
    terr = std::max(terr, std::fabs(double(Tw<63, 64>::im) + sin(2 * M_PI * 63 / 64)));
    terr = std::max(terr, std::fabs(ct::cos2pi(123, 2048) - cos(2 * M_PI * 123 / 2048)));
    terr = std::max(terr, std::fabs(ct::sin2pi(1999, 2048) - sin(2 * M_PI * 1999 / 2048)));
    printf("ct trig err %.3g\n", terr);
    if (terr > 6e-8) bad++;
#define REG(N) e = check_reg<N>(); printf("fft_reg<%d> err %.3g\n", N, e); if (e > 1e-6) bad++;
    REG(2) REG(4) REG(8) REG(16) REG(32) REG(64)
    for (int l = 1; l <= 12; ++l) {
        e = check_stockham(l);
        printf("stockham M=%d err %.3g\n", 1 << l, e);
        if (e > 2e-6) bad++;
    }
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
