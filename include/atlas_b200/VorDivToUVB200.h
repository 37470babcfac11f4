/*
 * VorDivToUVB200.h -- adaptor that plugs sptrans_vordiv_to_uv into atlas's VorDivToUV factory as type("b200"),
 * next to "local" (trans/local/VorDivToUVLocal.cc:25) and "ectrans".
 *
 * Interface implemented: trans::VorDivToUVImpl (ecmwf/atlas src/atlas/trans/VorDivToUV.h:34-58): truncation() and
 *     execute(nb_coeff, nb_fields, vorticity, divergence, U, V, config)
 * with the IFS-style spectral layout [m][n][re/im][field] on all four arrays, U = u cos(lat), V = v cos(lat)
 * (trans/local/VorDivToUVLocal.h:38-52).  Host or device pointers.  Registration, in one translation unit:
 *     static atlas::trans::VorDivToUVBuilder<atlas::trans::VorDivToUVB200> builder("b200");
 */
#pragma once

#include <sstream>
#include <string>

#include "sptrans_b200.h"

#include "atlas/functionspace/Spectral.h"
#include "atlas/runtime/Exception.h"
#include "atlas/trans/VorDivToUV.h"

namespace atlas {
namespace trans {

class VorDivToUVB200 : public VorDivToUVImpl {
public:
    VorDivToUVB200(const FunctionSpace& fs, const eckit::Configuration& config = util::NoConfig()):
        VorDivToUVB200(functionspace::Spectral(fs).truncation(), config) {}
    VorDivToUVB200(int truncation, const eckit::Configuration& config = util::NoConfig()): truncation_(truncation) {
        config.get("device", device_);
    }
    ~VorDivToUVB200() override = default;

    int truncation() const override { return truncation_; }

    void execute(const int nb_coeff, const int nb_fields, const double vorticity[], const double divergence[], double U[],
                 double V[], const eckit::Configuration& = util::NoConfig()) const override {
        // the reference asserts the coefficient count the same way (VorDivToUVLocal.cc:62-70)
        ATLAS_ASSERT(nb_coeff == (truncation_ + 1) * (truncation_ + 2) / 2, "nb_coeff does not match the truncation");
        const int rc = sptrans_vordiv_to_uv(truncation_, nb_fields, vorticity, divergence, U, V, device_);
        if (rc != SPTRANS_OK) {
            std::ostringstream msg;
            msg << "sptrans_b200: " << sptrans_last_error();
            throw_Exception(msg.str(), Here());
        }
    }

private:
    int truncation_;
    int device_{0};
};

}  // namespace trans
}  // namespace atlas
