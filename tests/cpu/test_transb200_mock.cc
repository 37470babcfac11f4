// Compiles include/atlas_b200/TransB200.h against the mock atlas headers (tests/cpu/mock_atlas) so that the
// adaptor's overrides are checked against the reference's TransImpl virtual signatures, registers it as
// type("b200") like trans/local/TransLocal.cc:57 registers "local", and drives it the way
// src/tests/trans/test_transgeneral.cc drives a Trans: build from a grid, invtrans of a unit coefficient.
// Exit codes: 0 ok (GPU present and result correct, or GPU absent and the refusal was loud), 1 failure.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "atlas_b200/TransB200.h"

namespace {
static atlas::trans::TransBuilderGrid<atlas::trans::TransB200> builder("b200", "b200");
}

int main() {
    using namespace atlas;
    const int N = 16, T = 15;
    auto d = std::make_shared<GridData>();
    d->name = "F16";
    d->gaussian = true;
    d->regular = true;
    d->nx.assign(2 * N, 4 * N);
    d->lat.resize(2 * N);
    std::vector<double> w(2 * N);
    sptrans_gaussian_latitudes(N, d->lat.data(), w.data());
    Grid grid(d);
    if (!trans::TransFactory::has("b200")) return 1;
    std::unique_ptr<const trans::TransImpl> t;
    try {
        t.reset(trans::TransFactory::build("b200", grid, T, util::NoConfig()));
    }
    catch (const eckit::Exception& e) {
        if (sptrans_device_count() == 0 && std::strstr(e.what(), "no CPU fallback")) {
            std::printf("no GPU: backend refused loudly (%s)\n", e.what());
            return 0;
        }
        std::printf("unexpected exception: %s\n", e.what());
        return 1;
    }
    if (t->type() != "b200" || t->truncation() != T || t->nb_spectral_coefficients() != size_t(T + 1) * (T + 2)) return 1;
    // coefficient (m=0,n=0) = 4  =>  every grid point = 4  (reference: test_trans.cc:343-420 test_nomesh)
    Field sp("sp", {idx_t(t->nb_spectral_coefficients())});
    Field gp("gp", {grid.size()});
    sp.data()[0] = 4.;
    t->invtrans(sp, gp);
    double err = 0;
    for (idx_t i = 0; i < grid.size(); ++i) err = std::fmax(err, std::fabs(gp.data()[i] - 4.));
    Field back("back", {idx_t(t->nb_spectral_coefficients())});
    t->dirtrans(gp, back);
    double err2 = std::fabs(back.data()[0] - 4.);
    for (size_t i = 1; i < t->nb_spectral_coefficients(); ++i) err2 = std::fmax(err2, std::fabs(back.data()[i]));
    // multi-level Fields in atlas's layouts: spectral (nspec2, nlev), grid-point (npts, nlev) level-fastest
    const idx_t nlev = 3;
    const idx_t nspec2 = idx_t(t->nb_spectral_coefficients());
    Field spl("spl", {nspec2, nlev}), gpl("gpl", {grid.size(), nlev}), backl("backl", {nspec2, nlev});
    for (idx_t l = 0; l < nlev; ++l) spl.data()[0 * nlev + l] = 1. + l;  // (m=0,n=0) of level l
    t->invtrans(spl, gpl);
    double err3 = 0;
    for (idx_t i = 0; i < grid.size(); ++i)
        for (idx_t l = 0; l < nlev; ++l) err3 = std::fmax(err3, std::fabs(gpl.data()[i * nlev + l] - (1. + l)));
    t->dirtrans(gpl, backl);
    for (idx_t l = 0; l < nlev; ++l) err3 = std::fmax(err3, std::fabs(backl.data()[l] - (1. + l)));
    // wind Field (npts, nlev, 2) from zero vorticity / divergence is zero; the adjoint Field overloads run
    Field vor("vor", {nspec2, nlev}), dv("div", {nspec2, nlev}), wind("wind", {grid.size(), nlev, 2});
    wind.data()[5] = 7.;
    t->invtrans_vordiv2wind(vor, dv, wind);
    for (idx_t i = 0; i < grid.size() * nlev * 2; ++i) err3 = std::fmax(err3, std::fabs(wind.data()[i]));
    t->invtrans_adj(gpl, backl);
    t->dirtrans_adj(spl, gpl);
    // unstructured grid through the same factory: coefficient (0,0) = 4 => every point = 4; no direct transform
    auto du = std::make_shared<GridData>();
    du->name = "unstructured";
    du->structured = false;
    du->points = {PointLonLat(0., 10.), PointLonLat(33., -10.), PointLonLat(271.5, 0.), PointLonLat(5., 90.)};
    Grid ugrid(du);
    std::unique_ptr<const trans::TransImpl> tu(trans::TransFactory::build("b200", ugrid, T, util::NoConfig()));
    Field spu("spu", {nspec2}), gpu("gpu", {ugrid.size()});
    spu.data()[0] = 4.;
    tu->invtrans(spu, gpu);
    for (idx_t i = 0; i < ugrid.size(); ++i) err3 = std::fmax(err3, std::fabs(gpu.data()[i] - 4.));
    bool threw_u = false;
    try {
        tu->dirtrans(gpu, spu);
    }
    catch (const eckit::NotImplemented&) {
        threw_u = true;
    }
    if (!threw_u) err3 = 1.;
    bool threw = false;
    try {
        t->dirtrans_wind2vordiv_adj(vor, dv, wind);
    }
    catch (const eckit::NotImplemented&) {
        threw = true;
    }
    std::printf("TransB200 via factory: invtrans err %.3e, dirtrans err %.3e, multi-level Field err %.3e, NotImplemented thrown: %d\n",
                err, err2, err3, threw);
    return (err < 1e-13 && err2 < 1e-13 && err3 < 1e-13 && threw) ? 0 : 1;
}
