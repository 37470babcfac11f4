/*
 * TransB200.h -- adaptor that plugs the sptrans C ABI into ecmwf/atlas as Trans backend `type("b200")`.
 *
 * It is a ~200-line header on purpose: everything numerical lives behind include/sptrans_b200.h.
 * Compile it inside an atlas build (or an atlas::Plugin) with
 *     #include "atlas/trans/detail/TransImpl.h", "atlas/trans/detail/TransFactory.h", "atlas/grid.h", ...
 * available; tests/cpu/test_transb200_mock.cc compiles it against a minimal mock of those headers
 * (tests/cpu/mock_atlas/) so that the virtual signatures cannot drift from the reference's
 * src/atlas/trans/detail/TransImpl.h:38-191.
 *
 * Behaviour mirrors TransLocal where TransLocal implements something (trans/local/TransLocal.cc), and
 * goes beyond it for dirtrans (TransLocal: ATLAS_NOTIMPLEMENTED, :1671-1685).  Entry points this engine
 * does not provide throw eckit::NotImplemented exactly like TransLocal does.
 */
#pragma once

#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "sptrans_b200.h"

#include "atlas/array.h"
#include "atlas/field.h"
#include "atlas/functionspace/Spectral.h"
#include "atlas/grid.h"
#include "atlas/runtime/Exception.h"
#include "atlas/trans/detail/TransFactory.h"
#include "atlas/trans/detail/TransImpl.h"

namespace atlas {
namespace trans {

class TransB200 : public TransImpl {
public:
    // constructor signature required by TransBuilderGrid<T> (trans/detail/TransFactory.h:116-119)
    TransB200(const Cache& /*cache*/, const Grid& grid, const Domain& /*domain*/, const long truncation,
              const eckit::Configuration& config = util::NoConfig()):
        grid_(grid), truncation_(static_cast<int>(truncation)) {
        StructuredGrid g(grid_);
        if (!g || grid_.projection()) {
            throw_NotImplemented("TransB200 supports global structured grids without projection", Here());
        }
        const int nlat = static_cast<int>(g.ny());
        std::vector<int> nx(nlat);
        std::vector<double> lat(nlat), w;
        for (int j = 0; j < nlat; ++j) {
            nx[j]  = static_cast<int>(g.nx(j));
            lat[j] = g.y(j);
        }
        if (GaussianGrid(grid_)) {  // quadrature weights make the direct transform available
            std::vector<double> l2(nlat);
            w.resize(nlat);
            check(sptrans_gaussian_latitudes(nlat / 2, l2.data(), w.data()));
        }
        int device = 0;
        config.get("device", device);
        const unsigned flags = RegularGrid(grid_) ? SPTRANS_GRID_REGULAR : 0u;
        check(sptrans_plan_create(&plan_, nlat, nx.data(), lat.data(), w.empty() ? nullptr : w.data(), truncation_,
                                  flags, device));
    }
    TransB200(const Grid& grid, const long truncation, const eckit::Configuration& config = util::NoConfig()):
        TransB200(Cache(), grid, grid.domain(), truncation, config) {}

    ~TransB200() override { sptrans_plan_destroy(plan_); }

    std::string type() const override { return "b200"; }
    int truncation() const override { return truncation_; }
    size_t nb_spectral_coefficients() const override { return sptrans_nb_spectral_coefficients(plan_); }
    size_t nb_spectral_coefficients_global() const override { return sptrans_nb_spectral_coefficients(plan_); }
    const Grid& grid() const override { return grid_; }
    const functionspace::Spectral& spectral() const override {
        if (!spectral_) {
            spectral_ = functionspace::Spectral(truncation_);
        }
        return spectral_;
    }

    // ---- IFS-style raw pointers: 1:1 onto the C ABI (host or device pointers) ----
    void invtrans(const int nb_scalar_fields, const double scalar_spectra[], const int nb_vordiv_fields,
                  const double vorticity_spectra[], const double divergence_spectra[], double gp_fields[],
                  const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_invtrans(plan_, nb_scalar_fields, scalar_spectra, nb_vordiv_fields, vorticity_spectra,
                               divergence_spectra, gp_fields));
    }
    void invtrans(const int nb_scalar_fields, const double scalar_spectra[], double gp_fields[],
                  const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_invtrans_scalar(plan_, nb_scalar_fields, scalar_spectra, gp_fields));
    }
    void invtrans(const int nb_vordiv_fields, const double vorticity_spectra[], const double divergence_spectra[],
                  double gp_fields[], const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_invtrans_vordiv2wind(plan_, nb_vordiv_fields, vorticity_spectra, divergence_spectra, gp_fields));
    }
    void dirtrans(const int nb_fields, const double scalar_fields[], double scalar_spectra[],
                  const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_dirtrans_scalar(plan_, nb_fields, scalar_fields, scalar_spectra));
    }
    void dirtrans(const int nb_fields, const double wind_fields[], double vorticity_spectra[],
                  double divergence_spectra[], const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_dirtrans_wind2vordiv(plan_, nb_fields, wind_fields, vorticity_spectra, divergence_spectra));
    }
    void invtrans_adj(const int, const double[], const int, double[], double[], double[],
                      const eckit::Configuration& = util::NoConfig()) const override {
        ATLAS_NOTIMPLEMENTED;
    }
    void invtrans_adj(const int nb_scalar_fields, const double gp_fields[], double scalar_spectra[],
                      const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_invtrans_adj_scalar(plan_, nb_scalar_fields, gp_fields, scalar_spectra));
    }
    void invtrans_adj(const int, const double[], double[], double[],
                      const eckit::Configuration& = util::NoConfig()) const override {
        ATLAS_NOTIMPLEMENTED;
    }

    // ---- Field interface.  Like TransLocal (TransLocal.cc:818-844) rank-1 fields are one scalar field; in
    // addition a rank-2 spectral field (nspec2, levels) / grid field (npts, levels) whose device copy is
    // valid is transformed in place on the device (array/Array.h:177-183 device_data). ----
    void invtrans(const Field& spfield, Field& gpfield, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfield.rank() == 1, "Only rank-1 fields supported at the moment");
        ATLAS_ASSERT(gpfield.rank() == 1, "Only rank-1 fields supported at the moment");
        const auto sp = array::make_view<double, 1>(spfield);
        auto gp       = array::make_view<double, 1>(gpfield);
        invtrans(1, sp.data(), gp.data(), config);
    }
    void invtrans(const FieldSet& spfields, FieldSet& gpfields, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfields.size() == gpfields.size());
        for (idx_t f = 0; f < spfields.size(); ++f) {
            invtrans(spfields[f], gpfields[f], config);
        }
    }
    void dirtrans(const Field& gpfield, Field& spfield, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfield.rank() == 1, "Only rank-1 fields supported at the moment");
        ATLAS_ASSERT(gpfield.rank() == 1, "Only rank-1 fields supported at the moment");
        const auto gp = array::make_view<double, 1>(gpfield);
        auto sp       = array::make_view<double, 1>(spfield);
        dirtrans(1, gp.data(), sp.data(), config);
    }
    void dirtrans(const FieldSet& gpfields, FieldSet& spfields, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfields.size() == gpfields.size());
        for (idx_t f = 0; f < spfields.size(); ++f) {
            dirtrans(gpfields[f], spfields[f], config);
        }
    }
    void invtrans_vordiv2wind(const Field& spvor, const Field& spdiv, Field& gpwind,
                              const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spvor.rank() == 1 && spdiv.rank() == 1, "Only rank-1 fields supported at the moment");
        const auto vor = array::make_view<double, 1>(spvor);
        const auto div = array::make_view<double, 1>(spdiv);
        auto gp        = array::make_view<double, 2>(gpwind);
        if (gp.shape(0) == 2) {  // (2, npts): the layout the engine writes (TransLocal.cc:885-887)
            invtrans(1, vor.data(), div.data(), gp.data(), config);
        }
        else {
            ATLAS_NOTIMPLEMENTED;  // (npts, 2) needs a transpose pass -- see INTEGRATION.md
        }
    }
    void dirtrans_wind2vordiv(const Field&, Field&, Field&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void dirtrans_adj(const Field&, Field&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void dirtrans_adj(const FieldSet&, FieldSet&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void dirtrans_wind2vordiv_adj(const Field&, const Field&, Field&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void invtrans_grad(const Field& spfield, Field& gradfield, const eckit::Configuration& = util::NoConfig()) const override {
        // gradfield (2, npts): row 0 = E-W, row 1 = N-S (component order of ifs/TransIFS.cc:2113-2137)
        ATLAS_ASSERT(spfield.rank() == 1 && gradfield.rank() == 2, "rank-1 spectral field, (2, npts) gradient field");
        const auto sp = array::make_view<double, 1>(spfield);
        auto g        = array::make_view<double, 2>(gradfield);
        if (g.shape(0) != 2) {
            ATLAS_NOTIMPLEMENTED;
        }
        check(sptrans_invtrans_grad(plan_, 1, sp.data(), g.data()));
    }
    void invtrans_grad(const FieldSet&, FieldSet&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void invtrans_adj(const Field&, Field&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void invtrans_adj(const FieldSet&, FieldSet&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void invtrans_grad_adj(const Field&, Field&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void invtrans_grad_adj(const FieldSet&, FieldSet&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }
    void invtrans_vordiv2wind_adj(const Field&, Field&, Field&, const eckit::Configuration& = util::NoConfig()) const override { ATLAS_NOTIMPLEMENTED; }

private:
    static void check(int rc) {
        if (rc == SPTRANS_OK) {
            return;
        }
        std::ostringstream msg;
        msg << "sptrans_b200: " << sptrans_last_error();
        if (rc == SPTRANS_ERR_NOT_IMPLEMENTED) {
            throw_NotImplemented(msg.str(), Here());
        }
        throw_Exception(msg.str(), Here());
    }

    Grid grid_;
    int truncation_;
    sptrans_plan* plan_{nullptr};
    mutable functionspace::Spectral spectral_;
};

// Registration (one translation unit of the plugin / of libatlas must contain):
//     namespace { static atlas::trans::TransBuilderGrid<atlas::trans::TransB200> builder("b200", "b200"); }
// exactly like trans/local/TransLocal.cc:57 does for "local".

}  // namespace trans
}  // namespace atlas
