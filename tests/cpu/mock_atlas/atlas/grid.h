// Mock of the atlas Grid handles used by a Trans backend (grid/detail/grid/Structured.h:298-312).
#pragma once
#include <memory>
#include <string>
#include <vector>
namespace atlas {
using idx_t = int;
class Domain {
public:
    explicit Domain(bool global = true, double ymin = -90., double ymax = 90.): global_(global), ymin_(ymin), ymax_(ymax) {}
    bool global() const { return global_; }
    double ymin() const { return ymin_; }
    double ymax() const { return ymax_; }
private:
    bool global_;
    double ymin_, ymax_;
};
class RectangularDomain : public Domain {  // domain/Domain.h
public:
    RectangularDomain(const Domain& d): Domain(d) {}
    explicit operator bool() const { return true; }
};
class Projection {
public:
    explicit operator bool() const { return false; }
};
class PointLonLat {  // util/Point.h
public:
    PointLonLat(double lon, double lat): lon_(lon), lat_(lat) {}
    double lon() const { return lon_; }
    double lat() const { return lat_; }
private:
    double lon_, lat_;
};
struct GridData {
    std::string name;
    std::vector<int> nx;
    std::vector<double> lat;
    bool gaussian = false, regular = false;
    bool lonlat_global = false;       // RegularLonLatGrid(grid) holds (global regular lon-lat grid)
    std::string yspace = "gaussian";  // StructuredGrid::yspace().type()
    double ymin = -90., ymax = 90.;   // RectangularDomain of a regional grid
    bool structured = true;
    bool global = true;               // domain().global()
    std::vector<PointLonLat> lonlat;  // filled for regional structured grids (Grid::lonlat())
    std::vector<PointLonLat> points;  // UnstructuredGrid
    std::vector<double> xmin;         // first longitude of every row (default 0: global rows)
    std::vector<double> dx;           // longitude increment of every row (default 360 / nx)
    std::shared_ptr<GridData> cropped;  // what Grid(this grid, a non-global domain) yields (the mock does not crop itself)
};
class Grid {
public:
    Grid() = default;
    explicit Grid(std::shared_ptr<GridData> d): d_(std::move(d)) {}
    Grid(const Grid& g, const Domain& domain): d_(domain.global() || !g.d_->cropped ? g.d_ : g.d_->cropped) {}  // grid/Grid.h:83
    explicit operator bool() const { return bool(d_); }
    Projection projection() const { return Projection(); }
    Domain domain() const { return Domain(d_->global, d_->ymin, d_->ymax); }
    idx_t size() const {
        if (!d_->structured) return static_cast<idx_t>(d_->points.size());
        idx_t s = 0;
        for (int n : d_->nx) s += n;
        return s;
    }
    const std::vector<PointLonLat>& lonlat() const { return d_->structured ? d_->lonlat : d_->points; }  // grid/Grid.h
    std::string name() const { return d_->name; }
    std::shared_ptr<GridData> d_;
};
class StructuredGrid : public Grid {
public:
    StructuredGrid() = default;
    StructuredGrid(const Grid& g): Grid(g) {}
    explicit operator bool() const { return d_ && d_->structured; }
    idx_t ny() const { return static_cast<idx_t>(d_->nx.size()); }
    idx_t nx(idx_t j) const { return d_->nx[j]; }
    double y(idx_t j) const { return d_->lat[j]; }
    double x(idx_t i, idx_t j) const {  // grid/detail/grid/Structured.h:312: xmin + i * dx
        const double x0 = d_->xmin.empty() ? 0. : d_->xmin[j];
        const double dx = d_->dx.empty() ? 360. / d_->nx[j] : d_->dx[j];
        return x0 + i * dx;
    }
    struct YSpace {
        std::string t;
        const std::string& type() const { return t; }
    };
    YSpace yspace() const { return YSpace{d_->yspace}; }
};
class GaussianGrid : public StructuredGrid {
public:
    GaussianGrid(const Grid& g): StructuredGrid(g) {}
    explicit operator bool() const { return d_ && d_->gaussian && d_->global; }
    long N() const { return static_cast<long>(d_->nx.size() / 2); }
};
class RegularLonLatGrid : public StructuredGrid {
public:
    RegularLonLatGrid(const Grid& g): StructuredGrid(g) {}
    explicit operator bool() const { return d_ && d_->lonlat_global; }
};
class RegularGrid : public StructuredGrid {
public:
    RegularGrid(const Grid& g): StructuredGrid(g) {}
    explicit operator bool() const { return d_ && d_->regular; }
};
}  // namespace atlas
