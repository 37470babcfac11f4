/*
 * LegendreCacheCreatorB200.h -- `trans::LegendreCacheCreator(grid, truncation, option::type("b200"))`.
 *
 * Mirrors trans/local/LegendreCacheCreatorLocal.{h,cc} (ecmwf/atlas): same interface (trans/LegendreCacheCreator.h:34-49),
 * registered with `LegendreCacheCreatorBuilder<LegendreCacheCreatorB200> builder("b200")` like LegendreCacheCreatorLocal.cc:30.
 *
 *   uid()      the reference's identifier scheme (LegendreCacheCreatorLocal.cc:66-119), formed by sptrans_legendre_cache_uid;
 *              the 66 strings src/tests/trans/test_trans_localcache.cc:264-360 expects are reproduced (tests/test_uid.py).
 *              The identifier keeps the prefix "local": the tables this backend generates on the device are bit-identical
 *              to TransLocal's and are exported in TransLocal's blob layout (TransLocal.cc:592-647), so a cache written by
 *              either backend serves both -- which is exactly what a shared uid expresses.
 *   create()   builds a temporary TransB200 (device-side Legendre generation: 0.1 s at T1279, not minutes) and exports the
 *              blob into a LegendreCache / a file (LegendreCacheCreatorLocal.cc:137-146).
 */
#pragma once

#include <cmath>
#include <string>
#include <vector>

#include "sptrans_b200.h"

#include "atlas/grid.h"
#include "atlas/runtime/Exception.h"
#include "atlas/trans/Cache.h"
#include "atlas/trans/LegendreCacheCreator.h"
#include "atlas/util/Config.h"
#include "atlas_b200/TransB200.h"

namespace atlas {
namespace trans {

class LegendreCacheCreatorB200 : public LegendreCacheCreatorImpl {
public:
    LegendreCacheCreatorB200(const Grid& grid, int truncation, const eckit::Configuration& config = util::NoConfig()):
        grid_(grid), truncation_(truncation) {
        flt_ = config.getBool("flt", false);  // the only option that enters the identifier (LegendreCacheCreatorLocal.cc:60-67)
        device_ = 0;
        config.get("device", device_);
    }
    ~LegendreCacheCreatorB200() override = default;

    bool supported() const override {  // LegendreCacheCreatorLocal.cc:126-134
        return static_cast<bool>(StructuredGrid(grid_)) && !grid_.projection();
    }

    std::string uid() const override {
        if (!unique_identifier_.empty()) {
            return unique_identifier_;
        }
        StructuredGrid structured(grid_);
        int kind = SPTRANS_UID_OTHER, n = 0;
        double south = 0., north = 0.;
        std::vector<double> lat;
        auto near = [](double a, double b) { return std::fabs(a - b) <= 1e-12 * std::fmax(1., std::fabs(b)); };
        if (grid_.projection()) {
            throw_NotImplemented("LegendreCacheCreatorB200::uid: grids with a projection hash the whole grid (Grid::hash)", Here());
        }
        else if (GaussianGrid(grid_)) {  // same cache for any global Gaussian grid (:82-85)
            kind = SPTRANS_UID_GAUSSIAN;
            n    = static_cast<int>(GaussianGrid(grid_).N());
        }
        else if (RegularLonLatGrid(grid_)) {  // same cache for any global regular grid (:86-108)
            RegularLonLatGrid g(grid_);
            const double dy_2 = 90. / double(g.ny());
            n = static_cast<int>(g.ny());
            if (near(g.y(0), 90.) && near(g.y(g.ny() - 1), -90.)) {
                kind = SPTRANS_UID_LONLAT;
            }
            else if (near(g.y(0), 90. - dy_2) && near(g.y(g.ny() - 1), -90. + dy_2)) {
                kind = SPTRANS_UID_SHIFTED_LONLAT;
            }
        }
        else if (RegularGrid(grid_) && structured.yspace().type() == "linear") {  // regional regular grids (:109-116)
            RectangularDomain domain(grid_.domain());
            ATLAS_ASSERT(domain);
            kind  = SPTRANS_UID_REGIONAL;
            south = domain.ymin();
            north = domain.ymax();
            n     = static_cast<int>(structured.ny());
        }
        if (kind == SPTRANS_UID_OTHER) {  // give_up: hash of the row latitudes (:46-58, :70-73)
            ATLAS_ASSERT(structured);
            for (idx_t j = 0; j < structured.ny(); ++j) {
                lat.push_back(structured.y(j));
            }
        }
        char buf[256];
        const int len = sptrans_legendre_cache_uid(buf, sizeof(buf), "local", truncation_, kind, n, south, north,
                                                   static_cast<int>(lat.size()), lat.empty() ? nullptr : lat.data(), flt_ ? 1 : 0);
        if (len < 0) {
            throw_Exception(std::string("LegendreCacheCreatorB200::uid: ") + sptrans_last_error(), Here());
        }
        unique_identifier_.assign(buf, static_cast<size_t>(len));
        return unique_identifier_;
    }

    void create(const std::string& path) const override {  // :137-139
        util::Config config;
        config.set("device", device_);
        config.set("write_legendre", path);   // option::write_legendre(path)
        TransB200 tmp(grid_, truncation_, config);
    }

    Cache create() const override {  // :141-146
        util::Config config;
        config.set("device", device_);
        TransB200 tmp(grid_, truncation_, config);
        LegendreCache cache(tmp.legendre_cache_size());
        tmp.export_legendre_cache(const_cast<void*>(cache.legendre().data()));
        return cache;
    }

    size_t estimate() const override { return sptrans_legendre_cache_estimate(truncation_); }  // :148-150

private:
    const Grid grid_;
    const int truncation_;
    bool flt_;
    int device_;
    mutable std::string unique_identifier_;
};

// Registration: namespace { static atlas::trans::LegendreCacheCreatorBuilder<atlas::trans::LegendreCacheCreatorB200> b("b200"); }

}  // namespace trans
}  // namespace atlas
