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

#include <cmath>
#include <cstdio>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "sptrans_b200.h"

#include "atlas/array.h"
#include "atlas/field.h"
#include "atlas/functionspace/Spectral.h"
#include "atlas/functionspace/StructuredColumns.h"
#include "atlas/grid.h"
#include "atlas/runtime/Exception.h"
#include "atlas/trans/detail/TransFactory.h"
#include "atlas/trans/detail/TransImpl.h"

namespace atlas {
namespace trans {

class TransB200 : public TransImpl {
public:
    // constructor signature required by TransBuilderGrid<T> (trans/detail/TransFactory.h:116-119)
    TransB200(const Cache& cache, const Grid& grid, const Domain& domain, const long truncation,
              const eckit::Configuration& config = util::NoConfig()):
        // like TransLocal (TransLocal.cc:322-336): the transform's grid is the given grid cropped to the domain
        grid_(domain.global() ? grid : Grid(grid, domain)), gridGlobal_(grid), truncation_(static_cast<int>(truncation)) {
        create_plan(config);
        if (!plan_is_points_) {
            // Legendre cache, as TransLocal handles it (TransLocal.cc:608-647): a cache handed in replaces the
            // tables (same blob layout, sizes checked), option "write_legendre" writes them out
            try {
                if (cache.legendre()) {
                    check(sptrans_import_legendre_cache(plan_, cache.legendre().data(), cache.legendre().size()));
                }
                std::string path;
                if (config.get("write_legendre", path) && !path.empty()) {
                    write_legendre(path);
                }
            }
            catch (...) {
                sptrans_plan_destroy(plan_);
                plan_ = nullptr;
                throw;
            }
        }
    }
    TransB200(const Grid& grid, const long truncation, const eckit::Configuration& config = util::NoConfig()):
        TransB200(Cache(), grid, grid.domain(), truncation, config) {}
    // constructor signature required by TransBuilderFunctionSpace<T> (trans/detail/TransFactory.h:99-113):
    // Trans(gp_functionspace, sp_functionspace, config) with the key "b200(StructuredColumns,Spectral)"
    // (TransFactory.cc:206-211), like TransIFSStructuredColumns (trans/ifs/TransIFSStructuredColumns.cc:23-29,35-39).
    // This backend works on the whole grid of the function space (one GPU holds every point): a StructuredColumns
    // distributed over several MPI ranks is refused exactly like TransLocal refuses mpi::size() > 1 (TransLocal.cc:338-340).
    TransB200(const Cache& cache, const FunctionSpace& gp, const FunctionSpace& sp,
              const eckit::Configuration& config = util::NoConfig()):
        TransB200(cache, grid_of(gp), grid_of(gp).domain(), truncation_of(sp), config) {
        spectral_ = functionspace::Spectral(sp);
    }

    ~TransB200() override {
        if (multi_) {
            sptrans_multi_destroy(multi_);   // owns the per-device plans (plan_ is the one of device 0)
        }
        else {
            sptrans_plan_destroy(plan_);
        }
    }

private:
    static Grid grid_of(const FunctionSpace& gp) {
        functionspace::StructuredColumns sc(gp);
        if (!sc) {
            throw_NotImplemented("TransB200(gp, sp): the grid-point function space must be StructuredColumns", Here());
        }
        if (static_cast<size_t>(sc.sizeOwned()) != static_cast<size_t>(sc.grid().size())) {
            throw_NotImplemented("TransB200(gp, sp): the StructuredColumns must hold the whole grid on this rank "
                                 "(like TransLocal, this backend is not MPI-distributed)", Here());
        }
        return sc.grid();
    }
    static long truncation_of(const FunctionSpace& sp) {
        functionspace::Spectral s(sp);
        if (!s) {
            throw_Exception("TransB200(gp, sp): the spectral function space must be Spectral", Here());
        }
        return s.truncation();
    }
    void write_legendre(const std::string& path) const {
        std::vector<char> blob(sptrans_legendre_cache_size(plan_));
        check(sptrans_export_legendre_cache(plan_, blob.data()));
        std::FILE* f = std::fopen(path.c_str(), "wb");
        if (!f || std::fwrite(blob.data(), 1, blob.size(), f) != blob.size()) {
            if (f) {
                std::fclose(f);
            }
            throw_Exception("TransB200: cannot write Legendre cache file " + path, Here());
        }
        std::fclose(f);
    }
    void create_plan(const eckit::Configuration& config) {
        int device = 0;
        config.get("device", device);
        StructuredGrid g(grid_);
        if (g && !grid_.domain().global() && StructuredGrid(gridGlobal_) && gridGlobal_.domain().global() && !grid_.projection()) {
            // Regional grid that is a cropping of a global structured grid (TransLocal.cc:371-531): the global grid's
            // Legendre functions, per-latitude zonal truncation and row FFTs, then the copy-out of the crop's longitudes.
            StructuredGrid gg(gridGlobal_);
            const int nlat_g = static_cast<int>(gg.ny()), nlat_c = static_cast<int>(g.ny());
            std::vector<int> nx(nlat_g), nxc(nlat_c), jlon(nlat_c);
            std::vector<double> lat(nlat_g);
            for (int j = 0; j < nlat_g; ++j) {
                nx[j]  = static_cast<int>(gg.nx(j));
                lat[j] = gg.y(j);
            }
            int jlat_min = 0;  // :441-447
            for (int j = 0; j < nlat_g; ++j) {
                if (gg.y(j) > g.y(0)) {
                    ++jlat_min;
                }
            }
            auto wrap = [](double angle) {  // :493-499
                double r = std::fmod(angle, 360.);
                return r < 0. ? r + 360. : r;
            };
            for (int j = 0; j < nlat_c; ++j) {  // jlonMin_: global points of the row west of the crop's first point (:518-529)
                const double lonmin = wrap(g.x(0, j));
                const int jg        = j + jlat_min;
                int count           = 0;
                for (int i = 0; i < nx[jg]; ++i) {
                    if (gg.x(i, jg) < lonmin - 1e-9) {
                        ++count;
                    }
                }
                jlon[j] = count % nx[jg];
                nxc[j]  = static_cast<int>(g.nx(j));
            }
            const unsigned flags = RegularGrid(gridGlobal_) ? SPTRANS_GRID_REGULAR : 0u;
            check(sptrans_plan_create_cropped(&plan_, nlat_g, nx.data(), lat.data(), truncation_, flags, device, jlat_min, nlat_c,
                                              nxc.data(), jlon.data()));
            plan_is_cropped_ = true;
            return;
        }
        if (!g || !grid_.domain().global()) {
            // Unstructured grid: the transform is evaluated point by point (TransLocal.cc:740-770, :1289-1392).
            // Regional REGULAR grids that are not a cropping of a global grid take the same route, which is what TransLocal
            // does too (no FFT, no zonal truncation: `no_nest`, :397-407, :462-467).  One table row per distinct latitude.
            std::vector<double> lon, lat;
            lon.reserve(grid_.size());
            lat.reserve(grid_.size());
            for (const PointLonLat p : grid_.lonlat()) {
                lon.push_back(p.lon());
                lat.push_back(p.lat());
            }
            check(sptrans_plan_create_points(&plan_, lon.size(), lon.data(), lat.data(), truncation_, device));
            plan_is_points_ = true;
            return;
        }
        if (grid_.projection()) {
            throw_NotImplemented("TransB200 supports structured grids without projection", Here());
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
        const unsigned flags = RegularGrid(grid_) ? SPTRANS_GRID_REGULAR : 0u;
        // config "gpus" = N > 1: the transform is sharded over N devices of this process (zonal wavenumbers x latitude
        // bands, exchange over NVLink peer memory: sptrans_multi_*); scalar invtrans / dirtrans only in that mode
        int gpus = 1;
        config.get("gpus", gpus);
        if (gpus > 1) {
            std::vector<int> devices(gpus);
            for (int r = 0; r < gpus; ++r) {
                devices[r] = device + r;
            }
            check(sptrans_multi_create(&multi_, nlat, nx.data(), lat.data(), w.empty() ? nullptr : w.data(), truncation_, flags,
                                       gpus, devices.data()));
            plan_ = sptrans_multi_plan(multi_, 0);
            return;
        }
        check(sptrans_plan_create(&plan_, nlat, nx.data(), lat.data(), w.empty() ? nullptr : w.data(), truncation_,
                                  flags, device));
    }

public:
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
        if (multi_) {
            check(sptrans_multi_invtrans_scalar(multi_, nb_scalar_fields, scalar_spectra, gp_fields));
            return;
        }
        check(sptrans_invtrans_scalar(plan_, nb_scalar_fields, scalar_spectra, gp_fields));
    }
    void invtrans(const int nb_vordiv_fields, const double vorticity_spectra[], const double divergence_spectra[],
                  double gp_fields[], const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_invtrans_vordiv2wind(plan_, nb_vordiv_fields, vorticity_spectra, divergence_spectra, gp_fields));
    }
    void dirtrans(const int nb_fields, const double scalar_fields[], double scalar_spectra[],
                  const eckit::Configuration& = util::NoConfig()) const override {
        if (multi_) {
            check(sptrans_multi_dirtrans_scalar(multi_, nb_fields, scalar_fields, scalar_spectra));
            return;
        }
        check(sptrans_dirtrans_scalar(plan_, nb_fields, scalar_fields, scalar_spectra));
    }
    void dirtrans(const int nb_fields, const double wind_fields[], double vorticity_spectra[],
                  double divergence_spectra[], const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_dirtrans_wind2vordiv(plan_, nb_fields, wind_fields, vorticity_spectra, divergence_spectra));
    }
    void invtrans_adj(const int nb_scalar_fields, const double gp_fields[], const int nb_vordiv_fields,
                      double vorticity_spectra[], double divergence_spectra[], double scalar_spectra[],
                      const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_invtrans_adj(plan_, nb_scalar_fields, gp_fields, nb_vordiv_fields, vorticity_spectra,
                                   divergence_spectra, scalar_spectra));
    }
    void invtrans_adj(const int nb_scalar_fields, const double gp_fields[], double scalar_spectra[],
                      const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_invtrans_adj_scalar(plan_, nb_scalar_fields, gp_fields, scalar_spectra));
    }
    void invtrans_adj(const int nb_vordiv_fields, const double wind_fields[], double vorticity_spectra[],
                      double divergence_spectra[], const eckit::Configuration& = util::NoConfig()) const override {
        check(sptrans_invtrans_vordiv2wind_adj(plan_, nb_vordiv_fields, wind_fields, vorticity_spectra, divergence_spectra));
    }

    // ---- Field interface.  TransLocal accepts rank-1 fields only (one scalar field, TransLocal.cc:818-844) and the
    // (2, npts) / (npts, 2) wind field (:871-897).  This backend accepts those and, in addition, multi-level Fields in
    // atlas's own layouts -- spectral (nspec2, levels), grid-point (nodes, levels), wind / gradient (nodes, levels, 2),
    // last index fastest, owned nodes first (functionspace/detail/StructuredColumns.h:104-137) -- exactly what TransIFS
    // packs on the host (ifs/TransIFS.cc:610-709, :1392-1437, :2113-2137).  If a Field's device copy is allocated and
    // current (field/Field.h:191-202) the transform runs on it in place: array().device_data() (array/Array.h:177-183)
    // goes straight to the engine, the level-fastest <-> row repack is a device transpose, nothing touches the host. ----
    void invtrans(const Field& spfield, Field& gpfield, const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spfield, nb_spectral_coefficients(), "spectral field");
        ATLAS_ASSERT(levels(gpfield, grid_.size(), "grid-point field") == nlev);
        if (spfield.rank() == 1 && gpfield.rank() == 1) {
            check(sptrans_invtrans_scalar(plan_, 1, in(spfield), out(gpfield)));
        }
        else {
            check(sptrans_invtrans_field(plan_, nlev, in(spfield), out(gpfield)));
        }
    }
    void invtrans(const FieldSet& spfields, FieldSet& gpfields, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfields.size() == gpfields.size());
        for (idx_t f = 0; f < spfields.size(); ++f) {
            invtrans(spfields[f], gpfields[f], config);
        }
    }
    void dirtrans(const Field& gpfield, Field& spfield, const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spfield, nb_spectral_coefficients(), "spectral field");
        ATLAS_ASSERT(levels(gpfield, grid_.size(), "grid-point field") == nlev);
        if (spfield.rank() == 1 && gpfield.rank() == 1) {
            check(sptrans_dirtrans_scalar(plan_, 1, in(gpfield), out(spfield)));
        }
        else {
            check(sptrans_dirtrans_field(plan_, nlev, in(gpfield), out(spfield)));
        }
    }
    void dirtrans(const FieldSet& gpfields, FieldSet& spfields, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfields.size() == gpfields.size());
        for (idx_t f = 0; f < spfields.size(); ++f) {
            dirtrans(gpfields[f], spfields[f], config);
        }
    }
    void invtrans_vordiv2wind(const Field& spvor, const Field& spdiv, Field& gpwind,
                              const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spvor, nb_spectral_coefficients(), "vorticity field");
        ATLAS_ASSERT(levels(spdiv, nb_spectral_coefficients(), "divergence field") == nlev);
        if (rows_layout(gpwind, nlev)) {  // (2, npts): the layout the engine works in (TransLocal.cc:885-887)
            check(sptrans_invtrans_vordiv2wind(plan_, 1, in(spvor), in(spdiv), out(gpwind)));
        }
        else {  // (npts, 2) (TransLocal.cc:888-893) or (npts, levels, 2)
            check_components(gpwind, nlev, "wind field");
            check(sptrans_invtrans_vordiv2wind_field(plan_, nlev, in(spvor), in(spdiv), out(gpwind)));
        }
    }
    void dirtrans_wind2vordiv(const Field& gpwind, Field& spvor, Field& spdiv,
                              const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spvor, nb_spectral_coefficients(), "vorticity field");
        ATLAS_ASSERT(levels(spdiv, nb_spectral_coefficients(), "divergence field") == nlev);
        if (rows_layout(gpwind, nlev)) {
            check(sptrans_dirtrans_wind2vordiv(plan_, 1, in(gpwind), out(spvor), out(spdiv)));
        }
        else {
            check_components(gpwind, nlev, "wind field");
            check(sptrans_dirtrans_wind2vordiv_field(plan_, nlev, in(gpwind), out(spvor), out(spdiv)));
        }
    }
    void invtrans_grad(const Field& spfield, Field& gradfield, const eckit::Configuration& = util::NoConfig()) const override {
        // component 0 = E-W, 1 = N-S (ifs/TransIFS.cc:2113-2137)
        const int nlev = levels(spfield, nb_spectral_coefficients(), "spectral field");
        if (rows_layout(gradfield, nlev)) {
            check(sptrans_invtrans_grad(plan_, 1, in(spfield), out(gradfield)));
        }
        else {
            check_components(gradfield, nlev, "gradient field");
            check(sptrans_invtrans_grad_field(plan_, nlev, in(spfield), out(gradfield)));
        }
    }
    void invtrans_grad(const FieldSet& spfields, FieldSet& gradfields, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfields.size() == gradfields.size());
        for (idx_t f = 0; f < spfields.size(); ++f) {
            invtrans_grad(spfields[f], gradfields[f], config);
        }
    }
    // adjoints (ATLAS_NOTIMPLEMENTED in TransLocal, TransLocal.cc:899-929, :1647-1667)
    void invtrans_adj(const Field& gpfield, Field& spfield, const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spfield, nb_spectral_coefficients(), "spectral field");
        ATLAS_ASSERT(levels(gpfield, grid_.size(), "grid-point field") == nlev);
        check(sptrans_invtrans_adj_field(plan_, nlev, in(gpfield), out(spfield)));
    }
    void invtrans_adj(const FieldSet& gpfields, FieldSet& spfields, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfields.size() == gpfields.size());
        for (idx_t f = 0; f < spfields.size(); ++f) {
            invtrans_adj(gpfields[f], spfields[f], config);
        }
    }
    void invtrans_grad_adj(const Field& gradfield, Field& spfield, const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spfield, nb_spectral_coefficients(), "spectral field");
        if (rows_layout(gradfield, nlev)) {
            check(sptrans_invtrans_grad_adj(plan_, 1, in(gradfield), out(spfield)));
        }
        else {
            check_components(gradfield, nlev, "gradient field");
            check(sptrans_invtrans_grad_adj_field(plan_, nlev, in(gradfield), out(spfield)));
        }
    }
    void invtrans_grad_adj(const FieldSet& gradfields, FieldSet& spfields, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfields.size() == gradfields.size());
        for (idx_t f = 0; f < spfields.size(); ++f) {
            invtrans_grad_adj(gradfields[f], spfields[f], config);
        }
    }
    void invtrans_vordiv2wind_adj(const Field& gpwind, Field& spvor, Field& spdiv,
                                  const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spvor, nb_spectral_coefficients(), "vorticity field");
        ATLAS_ASSERT(levels(spdiv, nb_spectral_coefficients(), "divergence field") == nlev);
        if (rows_layout(gpwind, nlev)) {
            check(sptrans_invtrans_vordiv2wind_adj(plan_, 1, in(gpwind), out(spvor), out(spdiv)));
        }
        else {
            check_components(gpwind, nlev, "wind field");
            check(sptrans_invtrans_vordiv2wind_adj_field(plan_, nlev, in(gpwind), out(spvor), out(spdiv)));
        }
    }
    void dirtrans_adj(const Field& spfield, Field& gpfield, const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spfield, nb_spectral_coefficients(), "spectral field");
        ATLAS_ASSERT(levels(gpfield, grid_.size(), "grid-point field") == nlev);
        check(sptrans_dirtrans_adj_field(plan_, nlev, in(spfield), out(gpfield)));
    }
    void dirtrans_adj(const FieldSet& spfields, FieldSet& gpfields, const eckit::Configuration& config = util::NoConfig()) const override {
        ATLAS_ASSERT(spfields.size() == gpfields.size());
        for (idx_t f = 0; f < spfields.size(); ++f) {
            dirtrans_adj(spfields[f], gpfields[f], config);
        }
    }
    // adjoint of wind -> vor/div (ATLAS_NOTIMPLEMENTED in TransLocal, TransLocal.cc:1661-1667; the reference's
    // test_2level_adjoint_test_with_vortdiv, test_transgeneral.cc:1725-1818, runs it through TransIFS)
    void dirtrans_wind2vordiv_adj(const Field& spvor, const Field& spdiv, Field& gpwind,
                                  const eckit::Configuration& = util::NoConfig()) const override {
        const int nlev = levels(spvor, nb_spectral_coefficients(), "vorticity field");
        ATLAS_ASSERT(levels(spdiv, nb_spectral_coefficients(), "divergence field") == nlev);
        if (rows_layout(gpwind, nlev)) {
            check(sptrans_dirtrans_wind2vordiv_adj(plan_, 1, in(spvor), in(spdiv), out(gpwind)));
        }
        else {
            check_components(gpwind, nlev, "wind field");
            check(sptrans_dirtrans_wind2vordiv_adj_field(plan_, nlev, in(spvor), in(spdiv), out(gpwind)));
        }
    }
    // the engine's tables in the reference's cache layout, for LegendreCacheCreatorB200::create (TransLocal's
    // export_legendre_, TransLocal.cc:592-647)
    size_t legendre_cache_size() const { return plan_is_points_ ? 0 : sptrans_legendre_cache_size(plan_); }
    void export_legendre_cache(void* out) const { check(sptrans_export_legendre_cache(plan_, out)); }

private:
    // Pointer the engine reads: the device copy of the Field if it is allocated and current, else the host copy.
    static const double* in(const Field& f) {
        if (f.deviceAllocated() && !f.deviceNeedsUpdate()) {
            return f.array().template device_data<double>();
        }
        return f.array().template host_data<double>();
    }
    // Pointer the engine writes; the other copy of the Field is marked stale.
    static double* out(Field& f) {
        if (f.deviceAllocated() && !f.deviceNeedsUpdate()) {
            f.setHostNeedsUpdate(true);
            return f.array().template device_data<double>();
        }
        f.setDeviceNeedsUpdate(true);
        return f.array().template host_data<double>();
    }
    // levels of a spectral (nspec2[, levels]) or scalar grid-point (nodes[, levels]) Field; like TransLocal a grid-point
    // Field may carry halo nodes after the owned ones (TransLocal.cc:825-830)
    static int levels(const Field& f, size_t min_leading, const char* what) {
        ATLAS_ASSERT(f.rank() == 1 || f.rank() == 2, what);
        ATLAS_ASSERT(static_cast<size_t>(f.shape()[0]) >= min_leading, what);
        return f.rank() == 1 ? 1 : static_cast<int>(f.shape()[1]);
    }
    // (2, npts): one level in the engine's own row layout
    bool rows_layout(const Field& f, int nlev) const {
        return nlev == 1 && f.rank() == 2 && f.shape()[0] == 2 && f.shape()[1] == grid_.size();
    }
    // (npts, 2) for one level, (npts, levels, 2) otherwise
    void check_components(const Field& f, int nlev, const char* what) const {
        const bool ok = (f.rank() == 2 && nlev == 1 && f.shape()[0] >= grid_.size() && f.shape()[1] == 2) ||
                        (f.rank() == 3 && f.shape()[0] >= grid_.size() && f.shape()[1] == nlev && f.shape()[2] == 2);
        if (!ok) {
            throw_NotImplemented(std::string("TransB200: unsupported shape of the ") + what, Here());
        }
    }
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
    Grid gridGlobal_;   // the grid the constructor was given (TransLocal::gridGlobal_, TransLocal.h:213); grid_ = it cropped to the domain
    int truncation_;
    sptrans_plan* plan_{nullptr};
    sptrans_multi* multi_{nullptr};   // config "gpus" > 1: the sharded plans of all devices (every other entry point then
                                      // reports "whole-transform entry points need an unsharded plan" from the C ABI)
    bool plan_is_points_{false};
    bool plan_is_cropped_{false};
    mutable functionspace::Spectral spectral_;
};

// Registration (one translation unit of the plugin / of libatlas must contain):
//     namespace {
//     static atlas::trans::TransBuilderGrid<atlas::trans::TransB200> builder("b200", "b200");
//     static atlas::trans::TransBuilderFunctionSpace<atlas::trans::TransB200> builder_fs("b200(StructuredColumns,Spectral)", "b200");
//     }
// exactly like trans/local/TransLocal.cc:57 does for "local" and trans/ifs/TransIFSStructuredColumns.cc:35-39 for "ectrans".

}  // namespace trans
}  // namespace atlas
