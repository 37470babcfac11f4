// Mock of the function-space handles a Trans backend sees: FunctionSpace (functionspace/FunctionSpace.h) carrying either a
// spectral truncation or a structured grid, functionspace::Spectral (functionspace/Spectral.h:43-61).
#pragma once
#include <cstddef>
#include <memory>
#include <string>
#include "atlas/grid.h"
namespace atlas {
class FunctionSpace {
public:
    FunctionSpace() = default;
    explicit FunctionSpace(int truncation): type_("Spectral"), t_(truncation) {}
    FunctionSpace(const Grid& grid, idx_t owned): type_("StructuredColumns"), grid_(grid), owned_(owned) {}
    const std::string& type() const { return type_; }
    int spectral_truncation() const { return t_; }
    const Grid& grid() const { return grid_; }
    idx_t owned() const { return owned_; }
private:
    std::string type_;
    int t_ = -1;
    Grid grid_;
    idx_t owned_ = 0;
};
namespace functionspace {
class Spectral {
public:
    Spectral() = default;
    explicit Spectral(int truncation): t_(truncation), set_(true) {}
    explicit Spectral(const FunctionSpace& fs): t_(fs.spectral_truncation()), set_(fs.type() == "Spectral" && fs.spectral_truncation() >= 0) {}
    explicit operator bool() const { return set_; }
    int truncation() const { return t_; }
    size_t nb_spectral_coefficients() const { return size_t(t_ + 1) * (t_ + 2); }
    static std::string type() { return "Spectral"; }
private:
    int t_ = -1;
    bool set_ = false;
};
}  // namespace functionspace
}  // namespace atlas
