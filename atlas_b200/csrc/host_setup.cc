// Host-side plan geometry for the B200 spectral-transform engine: everything the TransLocal
// constructor derives from the grid (ecmwf/atlas src/atlas/trans/local/TransLocal.cc:371-606)
// plus the per-latitude seeds of the device Legendre recurrence.
//
// Compiled WITHOUT fp contraction / -march flags: the seed columns must carry exactly the
// roundings of trans/local/LegendrePolynomials.cc so that the device-generated tables are
// bit-identical to the reference's.
#include <algorithm>
#include <cmath>
#include <limits>
#include <cstdio>
#include <numeric>
#include <sstream>
#include <thread>

#include "plan.hpp"

namespace sptrans {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

// Zonal truncation at one latitude row: linear / quadratic / cubic octahedral rule.
// (reference: trans/local/TransLocal.cc:272-300)
int fourier_truncation(int truncation, int nx, int /*nxmax*/, int ndgl, double lat, bool fullgrid) {
    const int t_linear = ndgl - 1;
    const int t_quadratic = ndgl * 2 / 3 - 1;
    int kept;
    if (fullgrid || truncation >= t_linear) {
        kept = (nx - 1) / 2;
    }
    else {
        const double c2 = std::pow(std::cos(lat), 2);
        if (truncation >= t_quadratic) {
            const double wq = 3 * (t_linear - truncation) / ndgl;  // integer quotient on purpose (reference :287)
            kept = static_cast<int>((nx - 1) / (2 + wq * c2));
        }
        else {
            kept = static_cast<int>((nx - 1) / (2 + c2) - 1);
        }
    }
    return std::min(truncation, kept);
}

// Gauss-Legendre nodes/weights on [-1,1] by Newton iteration on the three-term recurrence,
// returned as latitudes in degrees (north->south) and weights normalised to sum 1 -- the
// normalisation atlas uses (grid/detail/spacing/gaussian/Latitudes.cc:139-166: sum over 2N rows == 1).
void gaussian_quadrature(int N, double* lat_deg, double* weights) {
    const int n = 2 * N;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < N; ++i) {
        long double x = cosl(pi * (i + 0.75L) / (n + 0.5L));  // Tricomi-style first guess
        long double dp = 1;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1, p1 = x;
            for (int k = 2; k <= n; ++k) {
                long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
                p0 = p1;
                p1 = pk;
            }
            dp = n * (x * p1 - p0) / (x * x - 1);
            long double dx = p1 / dp;
            x -= dx;
            if (fabsl(dx) < 1e-19L) break;
        }
        // final derivative at the converged node
        {
            long double p0 = 1, p1 = x;
            for (int k = 2; k <= n; ++k) {
                long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
                p0 = p1;
                p1 = pk;
            }
            dp = n * (x * p1 - p0) / (x * x - 1);
        }
        long double w = 2 / ((1 - x * x) * dp * dp);  // sums to 2 over all nodes
        long double lat = asinl(x) * 180 / pi;
        lat_deg[i] = static_cast<double>(lat);
        lat_deg[n - 1 - i] = -static_cast<double>(lat);
        weights[i] = weights[n - 1 - i] = static_cast<double>(w / 2);
    }
}

int build_geometry(HostGeom& g, int nlat, const int* nx, const double* lat_deg, const double* weights, int T,
                   bool regular, int rank, int nranks) {
    if (nlat < 2 || T < 0 || !nx || !lat_deg || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("sptrans_plan_create: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    g.T = T;
    g.nlat = nlat;
    g.nleg = (nlat + 1) / 2;
    g.regular = regular;
    g.has_equator = (nlat % 2) == 1;
    g.nx.assign(nx, nx + nlat);
    g.lat_deg.assign(lat_deg, lat_deg + nlat);
    if (weights) g.weights.assign(weights, weights + nlat);
    g.rowoff.assign(nlat + 1, 0);
    g.nxmax = 0;
    for (int j = 0; j < nlat; ++j) {
        if (nx[j] < 1) {
            set_error("sptrans_plan_create: nx must be positive");
            return SPTRANS_ERR_INVALID;
        }
        g.rowoff[j + 1] = g.rowoff[j] + nx[j];
        g.nxmax = std::max(g.nxmax, nx[j]);
    }
    g.npts = g.rowoff[nlat];
    for (int j = 0; j < nlat; ++j) {
        const int mj = nlat - 1 - j;
        if (j > 0 && !(lat_deg[j] < lat_deg[j - 1])) {
            set_error("sptrans_plan_create: latitudes must decrease monotonically (north to south)");
            return SPTRANS_ERR_INVALID;
        }
        if (std::abs(lat_deg[j] + lat_deg[mj]) > 1e-9 || nx[j] != nx[mj]) {
            set_error("sptrans_plan_create: only global grids symmetric about the equator are supported "
                      "(regional / cropped domains: ATLAS_NOTIMPLEMENTED in this backend)");
            return SPTRANS_ERR_NOT_IMPLEMENTED;
        }
    }
    // first northern row at which wavenumber m is resolved (reference :462-488, global domain)
    g.nlat0.assign(T + 2, g.nleg);
    {
        int done = -1;
        for (int j = 0; j < nlat / 2; ++j) {
            int keep = fourier_truncation(T, nx[j], g.nxmax, nlat, lat_deg[j] * (M_PI / 180.), regular);
            keep = std::max(done, keep);
            for (int m = done + 1; m <= keep; ++m) g.nlat0[m] = j;
            done = keep;
        }
    }
    g.mmax.assign(g.nleg, -1);
    for (int j = 0; j < g.nleg; ++j)
        for (int m = 0; m <= T; ++m)
            if (g.nlat0[m] <= j) g.mmax[j] = m;
    // sharding: zonal wavenumbers by cost-balanced greedy assignment, latitude pairs by cost-balanced contiguous bands
    g.rank = rank;
    g.nranks = nranks;
    g.my_m.clear();
    g.owner.assign(T + 1, 0);
    g.band.assign(nranks + 1, g.nleg);
    g.band[0] = 0;
    if (nranks == 1) {
        for (int m = 0; m <= T; ++m) g.my_m.push_back(m);
        g.pair_begin = 0;
        g.pair_end = g.nleg;
    }
    else {
        std::vector<int> order(T + 1);
        std::iota(order.begin(), order.end(), 0);
        auto cost = [&](int m) { return static_cast<double>(T + 2 - m) * std::max(0, g.nleg - g.nlat0[m]); };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
        std::vector<double> load(nranks, 0.);
        for (int m : order) {
            int best = 0;
            for (int r = 1; r < nranks; ++r)
                if (load[r] < load[best]) best = r;
            load[best] += cost(m);
            g.owner[m] = best;
            if (best == rank) g.my_m.push_back(m);
        }
        std::sort(g.my_m.begin(), g.my_m.end());
        // contiguous bands of latitude pairs with ~equal Fourier-stage cost: a row pair of length n with zonal
        // wavenumbers up to L is one chirp-z transform of length ~ n + 2L (fourier.cu), i.e. ~ M log M work, plus a
        // per-row constant (block set-up)
        // Rows whose convolution does not fit the register-tiled kernels (n + 2L > 32 * 256) run on the shared-memory-pass
        // kernels or, beyond the single-CTA limit, on the row-mode kernels: 1.85x the time per unit of work, measured per rank
        // at TCo2559 on 8 GPUs (profiles/bench_tco2559_r02_n8_inv.json: bands of equal nominal cost took 6.6 ... 13.2 ms)
        auto fcost = [&](int j) {
            const double M = nx[j] + 2.0 * std::max(0, g.mmax[j]);
            return (M * std::log2(M + 2.0) + 2500.) * (M > 8192. ? 1.85 : 1.0);
        };
        double total = 0;
        for (int j = 0; j < g.nleg; ++j) total += fcost(j);
        std::vector<int> bound(nranks + 1, g.nleg);
        bound[0] = 0;
        double acc = 0;
        int r = 1;
        for (int j = 0; j < g.nleg && r < nranks; ++j) {
            acc += fcost(j);
            while (r < nranks && acc * nranks >= total * r) bound[r++] = j + 1;
        }
        g.band = bound;
        g.pair_begin = bound[rank];
        g.pair_end = bound[rank + 1];
    }
    // Legendre table blocks (m, parity): K rows (n ascending) x pitch latitudes
    g.tab_off.assign(2 * (T + 1), 0);
    g.tab_K.assign(2 * (T + 1), 0);
    g.tab_pitch.assign(T + 1, 0);
    g.sp_rowoff.assign(2 * (T + 1) + 1, 0);
    g.fb_rowoff.assign(T + 2, 0);
    long long off = 0;
    for (int m = 0; m <= T; ++m) {
        const int ncol = g.nleg - g.nlat0[m];
        g.tab_pitch[m] = round_up(std::max(ncol, 0), kBK > 16 ? kBK : 16);  // whole contraction steps of the direct GEMM
        for (int p = 0; p < 2; ++p) {
            const int K = num_n(T + 1, m, p);
            g.tab_K[2 * m + p] = K;
            // a sharded plan holds only the table blocks of its own zonal wavenumbers
            const bool mine = (nranks == 1) || (g.owner[m] == rank);
            g.tab_off[2 * m + p] = mine ? off : -1;
            // rows padded to the GEMM tile height so that tile loads never leave the block
            if (mine) off += static_cast<long long>(round_up(std::max(K, 1), kBK)) * g.tab_pitch[m];
            g.sp_rowoff[2 * m + p + 1] = g.sp_rowoff[2 * m + p] + round_up(std::max(K, 1), kBK);
        }
        g.fb_rowoff[m + 1] = g.fb_rowoff[m] + 2LL * std::max(ncol, 0);
    }
    g.tab_size = off;
    return SPTRANS_OK;
}

void set_io_layout(HostGeom& g, bool local_io) {
    g.local_io = local_io && g.nranks > 1;
    const int T = g.T, nlat = g.nlat;
    g.gp_rowoff.assign(nlat, -1);
    g.spec_off.assign(T + 1, -1);
    if (g.cropped) {   // work array of the Fourier stage: the rows of the band only; spectra are global
        long long off = 0;
        for (int j = g.pair_begin; j < g.pair_end; ++j) {
            g.gp_rowoff[j] = off;
            off += g.nx[j];
        }
        for (int j = g.pair_end - 1; j >= g.pair_begin; --j) {
            const int js = nlat - 1 - j;
            if (js == j) continue;
            g.gp_rowoff[js] = off;
            off += g.nx[js];
        }
        g.gp_stride = off + (off & 1);
        for (int m = 0; m <= T; ++m) g.spec_off[m] = static_cast<long long>(2 * T + 3 - m) * m / 2;
        g.spec_ncoef = static_cast<long long>(T + 1) * (T + 2) / 2;
        return;
    }
    if (!g.local_io) {
        for (int j = 0; j < nlat; ++j) g.gp_rowoff[j] = g.rowoff[j];
        g.gp_stride = g.npts;
        for (int m = 0; m <= T; ++m) g.spec_off[m] = static_cast<long long>(2 * T + 3 - m) * m / 2;  // reference :970
        g.spec_ncoef = static_cast<long long>(T + 1) * (T + 2) / 2;
        return;
    }
    // grid: northern rows [pair_begin, pair_end) in order, then their southern mirrors north -> south (the equator row
    // of a grid with an odd number of rows is a northern row and has no mirror)
    long long off = 0;
    for (int j = g.pair_begin; j < g.pair_end; ++j) {
        g.gp_rowoff[j] = off;
        off += g.nx[j];
    }
    for (int j = g.pair_end - 1; j >= g.pair_begin; --j) {
        const int js = nlat - 1 - j;
        if (js == j) continue;
        g.gp_rowoff[js] = off;
        off += g.nx[js];
    }
    g.gp_stride = off + (off & 1);  // even: 16-byte row starts of every field for the staged loads of the direct kernel
    long long c = 0;
    for (int m : g.my_m) {   // ascending
        g.spec_off[m] = c;
        c += T - m + 1;
    }
    g.spec_ncoef = c;
}

// Segment lists of the transposition between the two shardings.  Both sides enumerate (m ascending, parity,
// latitude ascending) so that the packed buffers of sender and receiver line up without any metadata exchange.
void build_exchange(const HostGeom& g, ExchangeLayout& ex) {
    const int R = g.nranks, me = g.rank;
    ex.m_side.clear();
    ex.band_side.clear();
    ex.m_side_rows.assign(R, 0);
    ex.band_side_rows.assign(R, 0);
    auto add = [&](std::vector<ExSeg>& v, long long& cursor, int m, int b0, int b1, int peer) -> long long {
        const int n0 = g.nlat0[m];
        const int ncol = g.nleg - n0;
        const int lo = std::max(b0, n0), hi = b1;
        long long rows = 0;
        if (hi <= lo) return 0;
        for (int p = 0; p < 2; ++p) {
            ExSeg s{};
            s.fb_row = g.fb_rowoff[m] + static_cast<long long>(p) * ncol + (lo - n0);
            s.buf_row = cursor;
            s.nrows = hi - lo;
            s.peer = peer;
            v.push_back(s);
            cursor += s.nrows;
            rows += s.nrows;
        }
        return rows;
    };
    long long cur = 0;
    for (int d = 0; d < R; ++d)           // my zonal wavenumbers, restricted to rank d's latitude band
        for (int m : g.my_m) ex.m_side_rows[d] += add(ex.m_side, cur, m, g.band[d], g.band[d + 1], d);
    cur = 0;
    for (int s = 0; s < R; ++s)           // rank s's zonal wavenumbers, restricted to my latitude band
        for (int m = 0; m <= g.T; ++m)
            if (g.owner[m] == s) ex.band_side_rows[s] += add(ex.band_side, cur, m, g.band[me], g.band[me + 1], s);
}

// Seeds of the Legendre recurrence for each latitude: cos(theta), the m=0 and m=1 columns (cosine /
// sine series, Belousov eq. 19/21 with the IFS normalisation) and the sectoral diagonal.
// Operation order reproduces trans/local/LegendrePolynomials.cc:24-45 (coefficients) and :55-130.
void legendre_seeds(int trc, int nlats, const double* lats_rad, std::vector<double>& xcos,
                    std::vector<double>& col0, std::vector<double>& col1, std::vector<double>& diag) {
    const size_t W = static_cast<size_t>(trc) + 1;
    std::vector<double> coef(W * W, 0.);  // coef[n*W + k]
    coef[0] = 2.;
    for (int n = 1; n <= trc; ++n) {
        double top = coef[0];
        for (int g = 1; g <= n; ++g) top *= std::sqrt(1. - 0.25 / (g * g));
        coef[n * W + n] = top;
        const int odd = n % 2;
        for (int g = 2; g <= n - odd; g += 2) {
            const double num = ((g - 1.) * (2. * n - g + 2.));
            const double den = (g * (2. * n - g + 1.));
            coef[n * W + (n - g)] = coef[n * W + (n - g + 2)] * num / den;
        }
    }
    for (int n = 1; n <= trc; n += 2) coef[n * W] = 0.;  // reference :101 zeroes zfn(n,0) for odd n

    xcos.assign(nlats, 0.);
    col0.assign(static_cast<size_t>(nlats) * W, 0.);
    col1.assign(static_cast<size_t>(nlats) * W, 0.);
    diag.assign(static_cast<size_t>(nlats) * W, 0.);

    auto work = [&](int j0, int j1) {
        std::vector<double> vs(W), vc(W);
        for (int j = j0; j < j1; ++j) {
            const double theta = (M_PI_2 - lats_rad[j]);
            double x = std::cos(theta);
            volatile double s = std::sqrt(1. - x * x);
            for (int k = 1; k <= trc; ++k) {
                vs[k] = std::sin(k * theta);
                vc[k] = std::cos(k * theta);
            }
            double inv_s = 0.;
            if (std::abs(s) <= std::sqrt(std::numeric_limits<double>::epsilon())) {
                x = 1.;
                s = 0.;
            }
            else {
                inv_s = 1. / s;
            }
            double* c0 = &col0[static_cast<size_t>(j) * W];
            double* c1 = &col1[static_cast<size_t>(j) * W];
            double* dg = &diag[static_cast<size_t>(j) * W];
            c0[0] = 1.;
            for (int n = 1; n <= trc; ++n) {
                const double* cf = &coef[n * W];
                const int k0 = (n % 2 == 0) ? 2 : 1;
                double a = (n % 2 == 0) ? 0.5 * cf[0] : 0.;
                double b = 0.0;
                const double q = 1. / std::sqrt(n * (n + 1.));
                for (int k = k0; k <= n; k += 2) {
                    a = a + cf[k] * vc[k];
                    b = b + q * cf[k] * k * vs[k];
                }
                c0[n] = a;
                c1[n] = b;
            }
            // sectoral values P_n^n with underflow flush (reference :122-130)
            dg[0] = c0[0];
            if (trc >= 1) dg[1] = c1[1];
            const double tiny = inv_s * std::numeric_limits<double>::min();
            for (int n = 2; n <= trc; ++n) {
                const double sq = std::sqrt((2. * n + 1.) / (2. * n));
                double v = dg[n - 1] * s * sq;
                if (std::abs(v) < tiny) v = 0.0;
                dg[n] = v;
            }
            xcos[j] = x;
        }
    };
    unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 64u));
    nt = std::min<unsigned>(nt, std::max(1, nlats));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t) {
        int j0 = static_cast<int>(static_cast<long long>(nlats) * t / nt);
        int j1 = static_cast<int>(static_cast<long long>(nlats) * (t + 1) / nt);
        pool.emplace_back(work, j0, j1);
    }
    for (auto& th : pool) th.join();
}

}  // namespace sptrans

// ---- Legendre-cache unique identifier (LegendreCacheCreatorLocal::uid, trans/local/LegendreCacheCreatorLocal.cc:34-119) ----
// The reference hashes with eckit::MD5 (third party, not vendored): RFC 1321, restated here.  eckit's Hash::add(const
// char*) feeds strlen bytes, add(bool) one byte, add(long) the eight bytes of the value.
namespace sptrans {
namespace {
struct Md5 {
    uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    unsigned char buf[64];
    uint64_t len = 0;
    static uint32_t rol(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }
    void block(const unsigned char* p) {
        static const int s[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                                  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t w[16];
        for (int i = 0; i < 16; ++i)
            w[i] = static_cast<uint32_t>(p[4 * i]) | (static_cast<uint32_t>(p[4 * i + 1]) << 8) |
                   (static_cast<uint32_t>(p[4 * i + 2]) << 16) | (static_cast<uint32_t>(p[4 * i + 3]) << 24);
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
        for (int i = 0; i < 64; ++i) {
            uint32_t f;
            int g;
            if (i < 16) { f = (b & c) | (~b & d); g = i; }
            else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
            else { f = c ^ (b | ~d); g = (7 * i) & 15; }
            const uint32_t k = static_cast<uint32_t>(std::floor(std::fabs(std::sin(static_cast<double>(i + 1))) * 4294967296.0));
            const uint32_t t = d;
            d = c;
            c = b;
            b = b + rol(a + f + k + w[g], s[i]);
            a = t;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d;
    }
    void update(const void* data, size_t n) {
        const unsigned char* p = static_cast<const unsigned char*>(data);
        for (size_t i = 0; i < n; ++i) {
            buf[len++ & 63] = p[i];
            if ((len & 63) == 0) block(buf);
        }
    }
    std::string hexdigest() {
        const uint64_t bits = len * 8;
        const unsigned char one = 0x80, zero = 0;
        update(&one, 1);
        while ((len & 63) != 56) update(&zero, 1);
        unsigned char lb[8];
        for (int i = 0; i < 8; ++i) lb[i] = static_cast<unsigned char>(bits >> (8 * i));
        update(lb, 8);
        char out[33];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) std::snprintf(out + 8 * i + 2 * j, 3, "%02x", (h[i] >> (8 * j)) & 0xffu);
        return std::string(out, 32);
    }
};
}  // namespace

std::string legendre_cache_uid(const char* prefix, int truncation, int kind, int n_or_ny, double south, double north,
                               int nlat, const double* lat_deg, bool flt) {
    std::ostringstream stream;
    stream << prefix << "-T" << truncation << "-";
    switch (kind) {
        case SPTRANS_UID_GAUSSIAN: stream << "GaussianN" << n_or_ny; break;                     // :82-85
        case SPTRANS_UID_LONLAT: stream << "L" << "-ny" << n_or_ny; break;                      // :97-100
        case SPTRANS_UID_SHIFTED_LONLAT: stream << "S" << "-ny" << n_or_ny; break;              // :101-104
        case SPTRANS_UID_REGIONAL:                                                              // :109-116
            stream << "Regional" << "-south" << south << "-north" << north << "-ny" << n_or_ny;
            break;
        default: {                                                                              // give_up, :70-73, :46-58
            Md5 h;
            for (int j = 0; j < nlat; ++j) {
                const long v = std::lround(lat_deg[j] * 1.e8);
                h.update(&v, sizeof(v));
            }
            stream << "grid-" << h.hexdigest().substr(0, 10);
        }
    }
    Md5 o;                                                                                      // hash(config), :60-67
    o.update("flt", 3);
    const bool b = flt;
    o.update(&b, sizeof(b));
    stream << "-OPT" << o.hexdigest().substr(0, 10);
    return stream.str();
}
}  // namespace sptrans
