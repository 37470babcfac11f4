// Mock of the abstract backend interface, restating the virtual signatures of
// ecmwf/atlas src/atlas/trans/detail/TransImpl.h:38-191 (6 inspectors, 12 Field/FieldSet virtuals,
// 8 raw-pointer virtuals; all const, all with a trailing const eckit::Configuration&).
// If the adaptor's `override`s stop matching these, tests/test_transb200_mock.py fails to compile.
#pragma once
#include <cstddef>
#include <string>
#include "atlas/field.h"
#include "atlas/functionspace/Spectral.h"
#include "atlas/grid.h"
#include "eckit/config/Configuration.h"
namespace atlas {
namespace util {
inline const eckit::Configuration& NoConfig() {
    static eckit::Configuration c;
    return c;
}
struct Object {
    virtual ~Object() = default;
};
}  // namespace util
namespace trans {
using Cfg = eckit::Configuration;
class TransImpl : public util::Object {
public:
    virtual std::string type() const { return "wrong value"; }
    virtual ~TransImpl() = default;
    virtual int truncation() const = 0;
    virtual size_t nb_spectral_coefficients() const = 0;
    virtual size_t nb_spectral_coefficients_global() const = 0;
    virtual const Grid& grid() const = 0;
    virtual const functionspace::Spectral& spectral() const = 0;

    virtual void dirtrans(const Field& gp, Field& sp, const Cfg& = util::NoConfig()) const = 0;
    virtual void dirtrans(const FieldSet& gp, FieldSet& sp, const Cfg& = util::NoConfig()) const = 0;
    virtual void dirtrans_wind2vordiv(const Field& gpwind, Field& spvor, Field& spdiv, const Cfg& = util::NoConfig()) const = 0;
    virtual void dirtrans_adj(const Field& sp, Field& gp, const Cfg& = util::NoConfig()) const = 0;
    virtual void dirtrans_adj(const FieldSet& sp, FieldSet& gp, const Cfg& = util::NoConfig()) const = 0;
    virtual void dirtrans_wind2vordiv_adj(const Field& spvor, const Field& spdiv, Field& gpwind, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans(const Field& sp, Field& gp, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans(const FieldSet& sp, FieldSet& gp, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_grad(const Field& sp, Field& grad, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_grad(const FieldSet& sp, FieldSet& grad, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_vordiv2wind(const Field& spvor, const Field& spdiv, Field& gpwind, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_adj(const Field& gp, Field& sp, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_adj(const FieldSet& gp, FieldSet& sp, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_grad_adj(const Field& grad, Field& gp, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_grad_adj(const FieldSet& grad, FieldSet& sp, const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_vordiv2wind_adj(const Field& gpwind, Field& spvor, Field& spdiv, const Cfg& = util::NoConfig()) const = 0;

    virtual void invtrans(const int nb_scalar_fields, const double scalar_spectra[], const int nb_vordiv_fields,
                          const double vorticity_spectra[], const double divergence_spectra[], double gp_fields[],
                          const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans(const int nb_scalar_fields, const double scalar_spectra[], double gp_fields[],
                          const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans(const int nb_vordiv_fields, const double vorticity_spectra[], const double divergence_spectra[],
                          double gp_fields[], const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_adj(const int nb_scalar_fields, const double gp_fields[], const int nb_vordiv_fields,
                              double vorticity_spectra[], double divergence_spectra[], double scalar_spectra[],
                              const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_adj(const int nb_scalar_fields, const double gp_fields[], double scalar_spectra[],
                              const Cfg& = util::NoConfig()) const = 0;
    virtual void invtrans_adj(const int nb_vordiv_fields, const double wind_fields[], double vorticity_spectra[],
                              double divergence_spectra[], const Cfg& = util::NoConfig()) const = 0;
    virtual void dirtrans(const int nb_fields, const double scalar_fields[], double scalar_spectra[],
                          const Cfg& = util::NoConfig()) const = 0;
    virtual void dirtrans(const int nb_fields, const double wind_fields[], double vorticity_spectra[],
                          double divergence_spectra[], const Cfg& = util::NoConfig()) const = 0;
};
}  // namespace trans
}  // namespace atlas
