// Mock of atlas::Field / FieldSet (field/Field.h:64-206): rank, shape, contiguous double storage, and the host/device
// state interface of field/Field.h:191-202 + array/Array.h:168-183 (this mock has no device copy).
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "atlas/grid.h"
namespace atlas {
class Field {
public:
    Field() = default;
    Field(const std::string& name, std::vector<idx_t> shape): name_(name), shape_(std::move(shape)) {
        size_t n = 1;
        for (idx_t s : shape_) n *= s;
        store_ = std::make_shared<std::vector<double>>(n, 0.);
    }
    int rank() const { return static_cast<int>(shape_.size()); }
    const std::vector<idx_t>& shape() const { return shape_; }
    double* data() { return store_->data(); }
    const double* data() const { return store_->data(); }
    // array().host_data<T>() / device_data<T>()  (array/Array.h:168-183)
    struct ArrayRef {
        double* p;
        template <typename T> T* host_data() const { return p; }
        template <typename T> T* device_data() const { return nullptr; }
    };
    ArrayRef array() const { return ArrayRef{store_->data()}; }
    bool deviceAllocated() const { return false; }
    bool hostNeedsUpdate() const { return false; }
    bool deviceNeedsUpdate() const { return true; }
    void setHostNeedsUpdate(bool) const {}
    void setDeviceNeedsUpdate(bool) const {}
private:
    std::string name_;
    std::vector<idx_t> shape_;
    std::shared_ptr<std::vector<double>> store_;
};
class FieldSet {
public:
    idx_t size() const { return static_cast<idx_t>(f_.size()); }
    Field& operator[](idx_t i) { return f_[i]; }
    const Field& operator[](idx_t i) const { return f_[i]; }
    void add(const Field& f) { f_.push_back(f); }
private:
    std::vector<Field> f_;
};
}  // namespace atlas
