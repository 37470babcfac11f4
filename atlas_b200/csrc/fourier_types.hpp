// Types shared by the Fourier kernels (fourier.cu), the register-tiled chirp-z engine (fft2_core.cuh) and its
// CPU thread-emulation test (tests/cpu/test_fft2_emul.cc).
#pragma once

namespace sptrans {

struct PairMeta {
    long long chirp_off;  // A_u (2L+1 entries) then C_i (n entries), double2 units
    long long filt_off;   // filter spectrum, M entries, in the order the forward transform leaves its output, scaled by 1/M
    long long rowN;       // grid offset of the northern row
    long long rowS;       // grid offset of the southern row
    long long tw_off;     // v1: two-level twiddle table Wa (M/64+1) then Wb (64); v2: W1[t] = e^{-2 pi i t/M}, t < 256
    int n;                // row length
    int L;                // zonal truncation at this latitude (mmax[j]); -1: nothing resolved
    int M;                // convolution length >= n + 2L (v1: 5-smooth multiple of 8; v2: m1 * 256)
    int has_s;            // 0 for the equator row of a grid with an odd number of latitudes
    int F;                // fields transformed by one block (v1: together; v2: one after the other)
    int sched;            // v1: index into the per-class pass schedules (precomputed on the host)
    int m1;               // v2: radix of the block-level pass (M = m1 * 256); 0: this pair runs on the v1 kernels
    int mode;             // 0: north+south rows packed into one complex transform of length n
                          // 1: every row on its own, even/odd samples packed, complex length n/2 (rows too long for mode 0)
};

// Arguments of the v2 Fourier kernels (one struct so that the CPU emulation can call the same bodies).
struct Fft2Args {
    const PairMeta* meta;
    int nf;
    int F;                      // fields per block (one after the other)
    int mlimit;                 // inverse: highest zonal wavenumber that enters
    int nb_uv;                  // leading fields that are wind components (scaled by scale_lat)
    const long long* fb_rowoff;
    const int* nlat0;
    int nleg;
    const double2* twid;        // W1 tables
    const double2* t256;        // e^{-2 pi i q l/256}, index q*16 + l
    const double2* chirp;
    const double2* filt;
    const double* scale_lat;    // inverse: 1/cos(lat); direct: cos(lat)-type scaling of wind input (may be null if nb_uv == 0)
    const double* weights;      // direct: quadrature weights per latitude pair
    double2* fb;                // Legendre<->Fourier exchange buffer (read by the inverse, written by the direct kernel)
    double* gp;                 // grid fields (written by the inverse, read by the direct kernel)
    long long npts;
    int adjoint;
    int gp_aligned16;           // grid buffer is 16-byte aligned (cp.async staging of the direct kernel)
};

}  // namespace sptrans
