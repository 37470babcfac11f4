// Mock of the atlas Grid handles used by a Trans backend (grid/detail/grid/Structured.h:298-312).
#pragma once
#include <memory>
#include <string>
#include <vector>
namespace atlas {
using idx_t = int;
class Domain {
public:
    bool global() const { return true; }
};
class Projection {
public:
    explicit operator bool() const { return false; }
};
struct GridData {
    std::string name;
    std::vector<int> nx;
    std::vector<double> lat;
    bool gaussian = false, regular = false;
};
class Grid {
public:
    Grid() = default;
    explicit Grid(std::shared_ptr<GridData> d): d_(std::move(d)) {}
    explicit operator bool() const { return bool(d_); }
    Projection projection() const { return Projection(); }
    Domain domain() const { return Domain(); }
    idx_t size() const {
        idx_t s = 0;
        for (int n : d_->nx) s += n;
        return s;
    }
    std::string name() const { return d_->name; }
    std::shared_ptr<GridData> d_;
};
class StructuredGrid : public Grid {
public:
    StructuredGrid() = default;
    StructuredGrid(const Grid& g): Grid(g) {}
    idx_t ny() const { return static_cast<idx_t>(d_->nx.size()); }
    idx_t nx(idx_t j) const { return d_->nx[j]; }
    double y(idx_t j) const { return d_->lat[j]; }
};
class GaussianGrid : public StructuredGrid {
public:
    GaussianGrid(const Grid& g): StructuredGrid(g) {}
    explicit operator bool() const { return d_ && d_->gaussian; }
};
class RegularGrid : public StructuredGrid {
public:
    RegularGrid(const Grid& g): StructuredGrid(g) {}
    explicit operator bool() const { return d_ && d_->regular; }
};
}  // namespace atlas
