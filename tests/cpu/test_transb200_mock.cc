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

#include "atlas_b200/LegendreCacheCreatorB200.h"
#include "atlas_b200/TransB200.h"

namespace {
static atlas::trans::TransBuilderGrid<atlas::trans::TransB200> builder("b200", "b200");
// Trans(gp_functionspace, sp_functionspace, option::type("b200")): trans/detail/TransFactory.h:99-113, key TransFactory.cc:206-211
static atlas::trans::TransBuilderFunctionSpace<atlas::trans::TransB200> builder_fs("b200(StructuredColumns,Spectral)", "b200");
// LegendreCacheCreator(grid, truncation, option::type("b200")): trans/local/LegendreCacheCreatorLocal.cc:30
static atlas::trans::LegendreCacheCreatorBuilder<atlas::trans::LegendreCacheCreatorB200> builder_cache("b200");
}

// host-only part: unique identifiers of the cache creator for the grid kinds the reference distinguishes
// (trans/local/LegendreCacheCreatorLocal.cc:66-119; expected strings of src/tests/trans/test_trans_localcache.cc:264-360)
static int check_uids() {
    using namespace atlas;
    int bad = 0;
    auto expect = [&](const Grid& g, int T, const char* want) {
        std::unique_ptr<trans::LegendreCacheCreatorImpl> c(trans::LegendreCacheCreatorFactory::build("b200", g, T));
        const std::string got = c->uid();
        if (got != want || !c->supported() || c->estimate() != size_t(T) * T * T / 2 * 8) {
            std::printf("uid mismatch: got %s want %s\n", got.c_str(), want);
            ++bad;
        }
    };
    auto gauss = std::make_shared<GridData>();   // "O320": any global Gaussian grid with N = 320
    gauss->gaussian = true;
    gauss->nx.assign(640, 20);
    gauss->lat.assign(640, 0.);
    expect(Grid(gauss), 639, "local-T639-GaussianN320-OPT4189816c2e");
    auto ll = std::make_shared<GridData>();      // "L90": 181 rows from 90 to -90
    ll->regular = ll->lonlat_global = true;
    ll->yspace = "linear";
    ll->nx.assign(181, 360);
    for (int j = 0; j < 181; ++j) ll->lat.push_back(90. - j);
    expect(Grid(ll), 20, "local-T20-L-ny181-OPT4189816c2e");
    auto crop = std::make_shared<GridData>();    // "L90" cropped to latitudes [-20, 20]: regional regular grid, linear spacing
    crop->regular = true;
    crop->global = false;
    crop->yspace = "linear";
    crop->ymin = -20.;
    crop->ymax = 20.;
    crop->nx.assign(41, 21);
    for (int j = 0; j < 41; ++j) crop->lat.push_back(20. - j);
    expect(Grid(crop), 20, "local-T20-Regional-south-20-north20-ny41-OPT4189816c2e");
    crop->yspace = "gaussian";                   // not linear: the reference gives up and hashes the latitudes
    expect(Grid(crop), 20, "local-T20-grid-7824deccdf-OPT4189816c2e");
    return bad;
}

int main() {
    using namespace atlas;
    if (check_uids() != 0) return 1;
    if (!trans::LegendreCacheCreatorFactory::has("b200")) return 1;
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
    // adjoint of wind -> vor/div through the Field overload: zero spectra give zero wind
    wind.data()[7] = 3.;
    t->dirtrans_wind2vordiv_adj(vor, dv, wind);
    for (idx_t i = 0; i < grid.size() * nlev * 2; ++i) err3 = std::fmax(err3, std::fabs(wind.data()[i]));
    // Trans(gp_functionspace, sp_functionspace): same plan through the FunctionSpace builder
    FunctionSpace gpfs(grid, grid.size()), spfs(T);
    std::unique_ptr<const trans::TransImpl> tf(trans::TransFactory::build("b200", gpfs, spfs, util::NoConfig()));
    Field gp2("gp2", {grid.size()});
    tf->invtrans(sp, gp2);
    for (idx_t i = 0; i < grid.size(); ++i) err3 = std::fmax(err3, std::fabs(gp2.data()[i] - gp.data()[i]));
    bool threw = false;
    try {   // a StructuredColumns that holds only part of the grid on this rank is refused (TransLocal.cc:338-340 analogue)
        FunctionSpace part(grid, grid.size() / 2);
        std::unique_ptr<const trans::TransImpl> bad(trans::TransFactory::build("b200", part, spfs, util::NoConfig()));
    }
    catch (const eckit::NotImplemented&) {
        threw = true;
    }
    // LegendreCacheCreator("b200"): the blob it creates is what a Trans built FROM that cache reproduces bit for bit
    std::unique_ptr<trans::LegendreCacheCreatorImpl> creator(trans::LegendreCacheCreatorFactory::build("b200", grid, T));
    trans::Cache cache = creator->create();
    const bool cache_ok = cache.legendre() && cache.legendre().size() > 0;
    std::unique_ptr<const trans::TransImpl> tc(trans::TransFactory::build("b200", cache, grid, T, util::NoConfig()));
    Field gp3("gp3", {grid.size()});
    Field spr("spr", {nspec2});
    for (idx_t i = 0; i < nspec2; ++i) spr.data()[i] = 1. / (1. + i);
    spr.data()[1] = 0.;
    Field gp4("gp4", {grid.size()});
    t->invtrans(spr, gp3);
    tc->invtrans(spr, gp4);
    for (idx_t i = 0; i < grid.size(); ++i)
        if (gp3.data()[i] != gp4.data()[i]) err3 = 1.;
    // Trans(global_grid, domain, truncation) with a regional domain: the cropped plan (FFT path with the global grid's zonal
    // truncation, TransLocal.cc:371-531), not the point-by-point path.  Rows 14..17 of F16 (two either side of the equator),
    // longitudes -11.25 .. 11.25 degrees (5 points, the first two west of Greenwich: jlonMin wraps around).
    {
        auto dc = std::make_shared<GridData>();
        dc->name = "F16-cropped";
        dc->regular = true;
        dc->global = false;
        dc->nx.assign(4, 5);
        dc->lat.assign(d->lat.begin() + 14, d->lat.begin() + 18);
        dc->xmin.assign(4, -11.25);
        dc->dx.assign(4, 5.625);
        dc->ymin = dc->lat.back();
        dc->ymax = dc->lat.front();
        d->cropped = dc;
        Domain region(false, dc->ymin, dc->ymax);
        std::unique_ptr<const trans::TransImpl> tr(trans::TransFactory::build("b200", trans::Cache(), grid, T, util::NoConfig(), region));
        Grid cgrid(grid, region);
        if (tr->grid().size() != cgrid.size() || cgrid.size() != 20) err3 = 1.;
        Field gpc("gpc", {cgrid.size()});
        tr->invtrans(spr, gpc);
        // the same values as the global transform at the crop's points: global row 14 + r, longitude index (62 + i) % 64
        for (int r = 0; r < 4; ++r)
            for (int i = 0; i < 5; ++i) {
                const double want = gp3.data()[(14 + r) * 64 + (62 + i) % 64];
                err3 = std::fmax(err3, std::fabs(gpc.data()[r * 5 + i] - want));
            }
    }
    // config "gpus": the single-process multi-device plan behind the same raw-pointer calls (two ranks on device 0 here)
    {
        util::Config two;
        two.set("gpus", 2);
        two.set("device", 0);
        std::unique_ptr<const trans::TransImpl> tm;
        try {
            tm.reset(trans::TransFactory::build("b200", grid, T, two));
        }
        catch (const eckit::Exception&) {   // a box with a single GPU has no device 1: fine, the path is covered by pytest
        }
        if (tm) {
            std::vector<double> g5(grid.size());
            tm->invtrans(1, spr.data(), g5.data());
            for (idx_t i = 0; i < grid.size(); ++i) err3 = std::fmax(err3, std::fabs(g5[i] - gp3.data()[i]) > 1e-13 ? 1. : 0.);
        }
    }
    std::printf("TransB200 via factory: invtrans err %.3e, dirtrans err %.3e, multi-level Field err %.3e, partial function space refused: %d, "
                "cache ok: %d\n", err, err2, err3, threw, cache_ok);
    return (err < 1e-13 && err2 < 1e-13 && err3 < 1e-13 && threw && cache_ok) ? 0 : 1;
}
