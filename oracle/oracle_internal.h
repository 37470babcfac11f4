// TEST INFRASTRUCTURE ONLY -- shared declarations between the two oracle translation units.
#pragma once
#include <cstddef>
#include <vector>
namespace orc {
void gaussian_quadrature_npole_equator(int N, double* lats_deg, double* weights);
void legendre_zfn(int trc, double* zfn);
void legendre_lat(int trc, double lat, double* legpol, double* zfn, std::vector<double>& vsin,
                  std::vector<double>& vcos);
void legendre_tables(int trc, int nlats, const double* lats, double* leg_sym, double* leg_asym,
                     const size_t* start_sym, const size_t* start_asym, int nthreads);
size_t num_n(int truncation, int m, bool symmetric);
size_t add_padding(size_t n);
size_t legendre_size(size_t truncation);
}  // namespace orc
