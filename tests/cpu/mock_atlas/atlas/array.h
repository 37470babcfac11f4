// Mock of array::make_view over a Field's contiguous host storage (array/ArrayView.h).
#pragma once
#include <cstddef>
#include "atlas/field.h"
namespace atlas {
namespace array {
template <typename T, int Rank>
class View {
public:
    View(T* p, const std::vector<idx_t>& shape): p_(p), shape_(shape) {}
    T* data() const { return p_; }
    idx_t shape(int i) const { return shape_[i]; }
    size_t size() const {
        size_t s = 1;
        for (idx_t v : shape_) s *= v;
        return s;
    }
private:
    T* p_;
    std::vector<idx_t> shape_;
};
template <typename T, int Rank>
View<T, Rank> make_view(Field& f) {
    return View<T, Rank>(f.data(), f.shape());
}
template <typename T, int Rank>
View<const T, Rank> make_view(const Field& f) {
    return View<const T, Rank>(f.data(), f.shape());
}
}  // namespace array
}  // namespace atlas
