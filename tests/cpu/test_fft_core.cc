// CPU unit test of atlas_b200/csrc/fft_core.cuh: the exact index / phase algebra of the GPU Fourier
// kernels (in-place DIF/DIT passes, digit-reversed filter, chirp-z with M >= n + 2L) run sequentially
// on the host and compared with a naive O(n^2) DFT.  Built and run by tests/test_fft_core_cpu.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../atlas_b200/csrc/fft_core.cuh"

using namespace sptrans::fftc;

static double urand() { return rand() / (double)RAND_MAX - 0.5; }

struct Tables {
    int n, L, M, logM, Wn;
    std::vector<double2> W, A, C, Bh;
};

static Tables build(int n, int L) {
    Tables t;
    t.n = n;
    t.L = L;
    t.M = conv_length(n, L, &t.logM);
    t.Wn = 8192;
    if (t.M > t.Wn) t.Wn = t.M;
    t.W.resize(t.Wn);
    for (int k = 0; k < t.Wn; ++k) t.W[k] = make_double2(std::cos(-2 * M_PI * k / t.Wn), std::sin(-2 * M_PI * k / t.Wn));
    t.A.resize(2 * L + 1);
    for (int u = 0; u <= 2 * L; ++u) {
        double ang = M_PI * (double)chirp_residue(u, 0, n) / n;
        t.A[u] = make_double2(std::cos(ang), std::sin(ang));
    }
    t.C.resize(n);
    for (int i = 0; i < n; ++i) {
        double ang = M_PI * (double)chirp_residue(i, -2LL * L, n) / n;
        t.C[i] = make_double2(std::cos(ang), std::sin(ang));
    }
    std::vector<double2> b(padded_len(t.M), make_double2(0, 0));
    for (int k = -2 * L; k <= n - 1; ++k) {
        double ang = -M_PI * (double)chirp_residue(k, 0, n) / n;
        int idx = ((k % t.M) + t.M) % t.M;
        b[pad(idx)] = make_double2(std::cos(ang), std::sin(ang));
    }
    fft_dif_all(b.data(), 1, t.logM, t.W.data(), t.Wn, 0, 1);
    t.Bh.resize(t.M);
    for (int k = 0; k < t.M; ++k) t.Bh[k] = make_double2(b[pad(k)].x / t.M, b[pad(k)].y / t.M);
    return t;
}

static int check(int n, int L) {
    Tables t = build(n, L);
    // ---- inverse: z_i = sum_{m=-L..L} Z_m e^{+2 pi i m i/n}
    std::vector<double2> Z(2 * L + 1);
    for (auto& z : Z) z = make_double2(urand(), urand());
    std::vector<double2> X(padded_len(t.M), make_double2(0, 0));
    for (int u = 0; u <= 2 * L; ++u) X[pad(u)] = cmul(Z[u], t.A[u]);
    fft_dif_all(X.data(), 1, t.logM, t.W.data(), t.Wn, 0, 1);
    fft_dit_all<false>(X.data(), 1, t.logM, t.W.data(), t.Wn, t.Bh.data(), 0, 1);
    double err_inv = 0, nrm = 0;
    for (int i = 0; i < n; ++i) {
        double2 got = cmul(X[pad(i)], t.C[i]);
        double re = 0, im = 0;
        for (int u = 0; u <= 2 * L; ++u) {
            long long m = u - L;
            double ang = 2 * M_PI * (double)(((m * i) % n + n) % n) / n;
            re += Z[u].x * std::cos(ang) - Z[u].y * std::sin(ang);
            im += Z[u].x * std::sin(ang) + Z[u].y * std::cos(ang);
        }
        err_inv = std::fmax(err_inv, std::hypot(got.x - re, got.y - im));
        nrm = std::fmax(nrm, std::hypot(re, im));
    }
    // ---- forward: G_m = (1/n) sum_i z_i e^{-2 pi i m i/n}, m = -L..L
    std::vector<double2> z(n);
    for (auto& v : z) v = make_double2(urand(), urand());
    std::fill(X.begin(), X.end(), make_double2(0, 0));
    for (int i = 0; i < n; ++i) X[pad(i)] = cmulc(z[i], t.C[i]);
    fft_dif_all(X.data(), 1, t.logM, t.W.data(), t.Wn, 0, 1);
    fft_dit_all<true>(X.data(), 1, t.logM, t.W.data(), t.Wn, t.Bh.data(), 0, 1);
    double err_fwd = 0, nrm2 = 0;
    for (int u = 0; u <= 2 * L; ++u) {
        double2 got = cmulc(X[pad(u)], t.A[u]);
        got.x /= n;
        got.y /= n;
        long long m = u - L;
        double re = 0, im = 0;
        for (int i = 0; i < n; ++i) {
            double ang = -2 * M_PI * (double)(((m * i) % n + n) % n) / n;
            re += z[i].x * std::cos(ang) - z[i].y * std::sin(ang);
            im += z[i].x * std::sin(ang) + z[i].y * std::cos(ang);
        }
        re /= n;
        im /= n;
        err_fwd = std::fmax(err_fwd, std::hypot(got.x - re, got.y - im));
        nrm2 = std::fmax(nrm2, std::hypot(re, im));
    }
    printf("n=%5d L=%5d M=%5d  inv rel err %.3e   fwd rel err %.3e\n", n, L, t.M, err_inv / nrm, err_fwd / nrm2);
    return (err_inv / nrm < 1e-12 && err_fwd / nrm2 < 1e-12) ? 0 : 1;
}

// ---- mixed-radix engine: same chirp-z algebra with M = 2^a 3^b 5^c and two-level twiddles ----
static int check_g(int n, int L) {
    const int M = conv_length_smooth(n + 2 * L, 1 << 15);
    const ScheduleG sc = make_schedule_g(M);
    std::vector<double2> Wa(M / 64 + 1), Wb(64);
    for (int k = 0; k <= M / 64; ++k) Wa[k] = make_double2(std::cos(-2 * M_PI * (64.0 * k) / M), std::sin(-2 * M_PI * (64.0 * k) / M));
    for (int k = 0; k < 64; ++k) Wb[k] = make_double2(std::cos(-2 * M_PI * k / M), std::sin(-2 * M_PI * k / M));
    std::vector<double2> A(2 * L + 1), C(n), Bh(M);
    for (int u = 0; u <= 2 * L; ++u) {
        double ang = M_PI * (double)chirp_residue(u, 0, n) / n;
        A[u] = make_double2(std::cos(ang), std::sin(ang));
    }
    for (int i = 0; i < n; ++i) {
        double ang = M_PI * (double)chirp_residue(i, -2LL * L, n) / n;
        C[i] = make_double2(std::cos(ang), std::sin(ang));
    }
    std::vector<double2> b(M, make_double2(0, 0));
    for (int k = -2 * L; k <= n - 1; ++k) {
        double ang = -M_PI * (double)chirp_residue(k, 0, n) / n;
        b[swz(((k % M) + M) % M)] = make_double2(std::cos(ang), std::sin(ang));
    }
    fft_dif_g(b.data(), 1, M, sc, Wa.data(), Wb.data(), 0, 1);
    for (int k = 0; k < M; ++k) Bh[k] = make_double2(b[swz(k)].x / M, b[swz(k)].y / M);
    std::vector<double2> Z(2 * L + 1);
    for (auto& z : Z) z = make_double2(urand(), urand());
    std::vector<double2> X(M, make_double2(0, 0));
    for (int u = 0; u <= 2 * L; ++u) X[swz(u)] = cmul(Z[u], A[u]);
    fft_dif_g(X.data(), 1, M, sc, Wa.data(), Wb.data(), 0, 1);
    fft_dit_g<false>(X.data(), 1, M, sc, Wa.data(), Wb.data(), Bh.data(), 0, 1);
    double err_inv = 0, nrm = 0;
    for (int i = 0; i < n; i += (n > 2000 ? 7 : 1)) {
        double2 got = cmul(X[swz(i)], C[i]);
        double re = 0, im = 0;
        for (int u = 0; u <= 2 * L; ++u) {
            long long m = u - L;
            double ang = 2 * M_PI * (double)(((m * i) % n + n) % n) / n;
            re += Z[u].x * std::cos(ang) - Z[u].y * std::sin(ang);
            im += Z[u].x * std::sin(ang) + Z[u].y * std::cos(ang);
        }
        err_inv = std::fmax(err_inv, std::hypot(got.x - re, got.y - im));
        nrm = std::fmax(nrm, std::hypot(re, im));
    }
    std::vector<double2> z(n);
    for (auto& v : z) v = make_double2(urand(), urand());
    std::fill(X.begin(), X.end(), make_double2(0, 0));
    for (int i = 0; i < n; ++i) X[swz(i)] = cmulc(z[i], C[i]);
    fft_dif_g(X.data(), 1, M, sc, Wa.data(), Wb.data(), 0, 1);
    fft_dit_g<true>(X.data(), 1, M, sc, Wa.data(), Wb.data(), Bh.data(), 0, 1);
    double err_fwd = 0, nrm2 = 0;
    for (int u = 0; u <= 2 * L; u += (L > 500 ? 5 : 1)) {
        double2 got = cmulc(X[swz(u)], A[u]);
        got.x /= n;
        got.y /= n;
        long long m = u - L;
        double re = 0, im = 0;
        for (int i = 0; i < n; ++i) {
            double ang = -2 * M_PI * (double)(((m * i) % n + n) % n) / n;
            re += z[i].x * std::cos(ang) - z[i].y * std::sin(ang);
            im += z[i].x * std::sin(ang) + z[i].y * std::cos(ang);
        }
        re /= n;
        im /= n;
        err_fwd = std::fmax(err_fwd, std::hypot(got.x - re, got.y - im));
        nrm2 = std::fmax(nrm2, std::hypot(re, im));
    }
    // ---- fused boundary passes: generator input, sink output (what the kernels use) ----
    {
        std::vector<double2> Y(M, make_double2(-7, 9));  // garbage: the first pass must overwrite everything
        std::vector<double2> outz(n);
        auto gen = [&](int, int e) { return e <= 2 * L ? cmul(Z[e], A[e]) : make_double2(0, 0); };
        auto sink = [&](int, int i, double2 v) { if (i < n) outz[i] = cmul(v, C[i]); };
        fft_dif_g_gen(Y.data(), 1, M, sc, Wa.data(), Wb.data(), 0, 1, gen);
        fft_dif_g_rest(Y.data(), 1, M, sc, Wa.data(), Wb.data(), 0, 1);
        fft_dit_g_sink<false>(Y.data(), 1, M, sc, Wa.data(), Wb.data(), Bh.data(), 0, 1, sink);
        double e2 = 0;
        for (int i = 0; i < n; i += (n > 2000 ? 7 : 1)) {
            double re = 0, im = 0;
            for (int u = 0; u <= 2 * L; ++u) {
                long long m = u - L;
                double ang = 2 * M_PI * (double)(((m * i) % n + n) % n) / n;
                re += Z[u].x * std::cos(ang) - Z[u].y * std::sin(ang);
                im += Z[u].x * std::sin(ang) + Z[u].y * std::cos(ang);
            }
            e2 = std::fmax(e2, std::hypot(outz[i].x - re, outz[i].y - im));
        }
        err_inv = std::fmax(err_inv, e2);
    }
    printf("mixed n=%5d L=%5d M=%5d passes=%d [", n, L, M, sc.npass);
    for (int p = 0; p < sc.npass; ++p) printf("%d ", sc.radix[p]);
    printf("]  inv %.3e  fwd %.3e\n", err_inv / nrm, err_fwd / nrm2);
    return (err_inv / nrm < 1e-12 && err_fwd / nrm2 < 1e-12) ? 0 : 1;
}

// ---- direct transforms of lengths without prime factors above 23 (no chirp-z): forward leaves frequency k at dif_output_position(k), the
// inverse takes its input in that order.  n need not be a multiple of 8: the buffer has swz_len(n) slots. ----
static int check_direct(int n) {
    const ScheduleG sc = make_schedule_g(n);
    int prod = 1;
    for (int p = 0; p < sc.npass; ++p) prod *= sc.radix[p];
    if (prod != n || !is_direct_length(n)) {
        printf("direct n=%d: schedule does not cover the length\n", n);
        return 1;
    }
    const int PL = swz_len(n);
    std::vector<double2> Wa(n / 64 + 1), Wb(64);
    for (int k = 0; k <= n / 64; ++k) Wa[k] = make_double2(std::cos(-2 * M_PI * (64.0 * k) / n), std::sin(-2 * M_PI * (64.0 * k) / n));
    for (int k = 0; k < 64; ++k) Wb[k] = make_double2(std::cos(-2 * M_PI * k / n), std::sin(-2 * M_PI * k / n));
    std::vector<double2> z(n);
    for (auto& v : z) v = make_double2(urand(), urand());
    // two sequences back to back, like the kernels hold them
    std::vector<double2> X(2 * PL, make_double2(0, 0));
    for (int i = 0; i < n; ++i) X[swz(i)] = X[PL + swz(i)] = z[i];
    fft_dif_g<true>(X.data(), 2, n, sc, Wa.data(), Wb.data(), 0, 1, PL);   // two sequences, swz_len(n) slots apart
    double err_f = 0, nrm = 0;
    for (int k = 0; k < n; k += (n > 1500 ? 11 : 1)) {
        double re = 0, im = 0;
        for (int i = 0; i < n; ++i) {
            const double ang = -2 * M_PI * (double)(((long long)k * i) % n) / n;
            re += z[i].x * std::cos(ang) - z[i].y * std::sin(ang);
            im += z[i].x * std::sin(ang) + z[i].y * std::cos(ang);
        }
        const double2 got = X[PL + swz(dif_output_position(sc, n, k))];
        err_f = std::fmax(err_f, std::hypot(got.x - re, got.y - im));
        nrm = std::fmax(nrm, std::hypot(re, im));
    }
    // inverse of the forward = n * identity
    fft_dit_g<false, true>(X.data(), 1, n, sc, Wa.data(), Wb.data(), nullptr, 0, 1);
    double err_i = 0;
    for (int i = 0; i < n; ++i) err_i = std::fmax(err_i, std::hypot(X[swz(i)].x / n - z[i].x, X[swz(i)].y / n - z[i].y));
    printf("direct n=%5d passes=%d [", n, sc.npass);
    for (int p = 0; p < sc.npass; ++p) printf("%d ", sc.radix[p]);
    printf("]  fwd %.3e  inv %.3e\n", err_f / nrm, err_i);
    return (err_f / nrm < 1e-12 && err_i < 1e-12) ? 0 : 1;
}

// every length the plan may hand to the direct kernels (row lengths up to the single-CTA limit without a prime factor above 23):
// the schedule covers the length, fits the fixed-size pass arrays, and the digit reversal is a permutation
static int check_all_direct_schedules(int nmax) {
    int bad = 0, count = 0, most = 0;
    std::vector<char> seen;
    for (int n = 1; n <= nmax; ++n) {
        if (!is_direct_length(n)) continue;
        ++count;
        const ScheduleG sc = make_schedule_g(n);
        long long prod = 1;
        for (int p = 0; p < sc.npass; ++p) prod *= sc.radix[p];
        if (prod != n || sc.npass > 10) {
            printf("direct schedule n=%d: product %lld, %d passes\n", n, prod, sc.npass);
            ++bad;
            continue;
        }
        most = sc.npass > most ? sc.npass : most;
        if (n % 7 == 0 || n < 600) {   // (a subset keeps the test fast)
            seen.assign(n, 0);
            for (int k = 0; k < n; ++k) {
                const int pos = dif_output_position(sc, n, k);
                if (pos < 0 || pos >= n || seen[pos]) {
                    printf("direct schedule n=%d: output positions are not a permutation (k=%d)\n", n, k);
                    ++bad;
                    break;
                }
                seen[pos] = 1;
            }
        }
    }
    printf("direct schedules: %d lengths up to %d, at most %d passes, %d bad\n", count, nmax, most, bad);
    return bad;
}

int main() {
    int bad = check_all_direct_schedules(13824);
    const int direct[] = {17, 19, 23, 68, 76, 92, 4301, 7429, 4692, 5060, 4788, 3876, 391, 18, 30, 45, 63, 13, 26, 98, 1001, 77, 20, 24, 28, 36, 44, 52, 56, 84, 132, 144, 364, 572, 1092, 1456, 2184, 4004, 5096, 5120, 4732, 3432};
    for (int n : direct) bad += check_direct(n);
    const int cases[][2] = {{20, 9}, {24, 7}, {28, 0}, {144, 31}, {36, 17}, {1616, 399}, {5136, 1279}, {5132, 1279},
                            {2568, 1279}, {128, 31}, {9, 4}, {7, 3}, {4, 1}};
    for (auto& c : cases) bad += check(c[0], c[1]);
    const int cases_g[][2] = {{20, 9}, {24, 7}, {144, 31}, {1616, 399}, {5136, 1279}, {3000, 999}, {2052, 683}, {4000, 1279},
                              {10256, 2559}, {36, 17}, {100, 33}, {448, 148}, {1000, 332}, {2700, 899}, {3600, 1199}, {28, 0}};
    for (auto& c : cases_g) bad += check_g(c[0], c[1]);
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad;
}
