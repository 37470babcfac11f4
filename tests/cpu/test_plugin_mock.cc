// Compiles the plugin translation unit (plugin/atlas-b200/src/B200Plugin.cc) against the mock atlas headers and checks
// that its static initialisers register the plugin and the "b200" Trans backend, as loading the real plugin would.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../../plugin/atlas-b200/src/B200Plugin.cc"

int main() {
    const bool plugin = !atlas::Plugin::loaded().empty() && atlas::Plugin::loaded().front() == "atlas-b200";
    const bool backend = atlas::trans::TransFactory::has("b200");
    // VorDivToUV through its own factory: (zeta_1^0 = 1) -> U_0^0 = -eps(1,0) lap(1) ... = a / sqrt(3) (Temperton 1991, 2.12)
    bool vd2uv = atlas::trans::VorDivToUVFactory::has("b200");
    if (vd2uv) {
        const int T = 3, ncoef = (T + 1) * (T + 2) / 2;
        std::unique_ptr<atlas::trans::VorDivToUVImpl> op(atlas::trans::VorDivToUVFactory::build("b200", T));
        std::vector<double> vor(2 * ncoef, 0.), div(2 * ncoef, 0.), U(2 * ncoef, -1.), V(2 * ncoef, -1.);
        vor[2 * 1] = 1.;  // (m = 0, n = 1), real part
        try {
            op->execute(ncoef, 1, vor.data(), div.data(), U.data(), V.data());
            const double a = 6371229.;
            vd2uv = op->truncation() == T && std::fabs(U[0] - a / std::sqrt(3.)) < 1e-9 * a && std::fabs(V[0]) < 1e-9;
        }
        catch (const eckit::Exception& e) {
            vd2uv = sptrans_device_count() == 0 && std::strstr(e.what(), "no CPU fallback") != nullptr;
        }
    }
    std::printf("plugin registered: %d, backend registered: %d, VorDivToUV ok: %d\n", plugin, backend, vd2uv);
    return plugin && backend && vd2uv ? 0 : 1;
}
