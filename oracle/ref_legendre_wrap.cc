// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI wrapper around the *unmodified* reference source
//   /root/reference/src/atlas/trans/local/LegendrePolynomials.cc
// which oracle/Makefile compiles in place (nothing is copied into this repo)
// into oracle/_ref/libref_legendre.so.  It is used to pin the oracle's own
// restatement of the Legendre-polynomial code (oracle/sht_oracle.cc) and to
// generate the golden vectors under tests/golden/.
#include "atlas/trans/local/LegendrePolynomials.h"

extern "C" {

// compute_zfn, LegendrePolynomials.cc:24-45
void ref_compute_zfn(int trc, double* zfn) { atlas::trans::compute_zfn(trc, zfn); }

// compute_legendre_polynomials_lat, LegendrePolynomials.cc:47-151
void ref_legendre_lat(int trc, double lat_rad, double* legpol, double* zfn) {
    atlas::trans::LegendrePolynomialsWorkspace w(trc);
    atlas::trans::compute_legendre_polynomials_lat(trc, lat_rad, legpol, zfn, w);
}

// compute_legendre_polynomials, LegendrePolynomials.cc:154-209
void ref_legendre_tables(int trc, int nlats, const double* lats_rad, double* leg_sym, double* leg_asym,
                         size_t* start_sym, size_t* start_asym) {
    atlas::trans::compute_legendre_polynomials(trc, nlats, lats_rad, leg_sym, leg_asym, start_sym, start_asym);
}

}  // extern "C"
