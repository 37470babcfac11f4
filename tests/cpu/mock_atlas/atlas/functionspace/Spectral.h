// Mock of functionspace::Spectral (functionspace/Spectral.h:43-61): truncation holder.
#pragma once
namespace atlas {
namespace functionspace {
class Spectral {
public:
    Spectral() = default;
    explicit Spectral(int truncation): t_(truncation), set_(true) {}
    explicit operator bool() const { return set_; }
    int truncation() const { return t_; }
    size_t nb_spectral_coefficients() const { return size_t(t_ + 1) * (t_ + 2); }
private:
    int t_ = -1;
    bool set_ = false;
};
}  // namespace functionspace
}  // namespace atlas
