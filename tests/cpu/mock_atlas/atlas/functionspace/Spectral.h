// Mock of functionspace::Spectral (functionspace/Spectral.h:43-61): truncation holder.
#pragma once
#include <cstddef>
namespace atlas {
class FunctionSpace {  // functionspace/FunctionSpace.h handle; here it only carries a spectral truncation
public:
    FunctionSpace() = default;
    explicit FunctionSpace(int truncation): t_(truncation) {}
    int spectral_truncation() const { return t_; }
private:
    int t_ = -1;
};
namespace functionspace {
class Spectral {
public:
    Spectral() = default;
    explicit Spectral(int truncation): t_(truncation), set_(true) {}
    explicit Spectral(const FunctionSpace& fs): t_(fs.spectral_truncation()), set_(fs.spectral_truncation() >= 0) {}
    explicit operator bool() const { return set_; }
    int truncation() const { return t_; }
    size_t nb_spectral_coefficients() const { return size_t(t_ + 1) * (t_ + 2); }
private:
    int t_ = -1;
    bool set_ = false;
};
}  // namespace functionspace
}  // namespace atlas
