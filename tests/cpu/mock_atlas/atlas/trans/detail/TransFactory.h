// Mock of the backend registry (trans/detail/TransFactory.h:114-129, util/Factory.h:94-113).
#pragma once
#include <map>
#include <string>
#include "atlas/trans/detail/TransImpl.h"
namespace atlas {
namespace trans {
// trans/Cache.h:41-136: a Cache hands out raw byte entries; only the Legendre entry matters to a backend
class TransCacheEntry {
public:
    TransCacheEntry() = default;
    TransCacheEntry(const void* data, size_t size): data_(data), size_(size) {}
    operator bool() const { return size_ != 0; }
    size_t size() const { return size_; }
    const void* data() const { return data_; }
private:
    const void* data_ = nullptr;
    size_t size_ = 0;
};
class Cache {
public:
    Cache() = default;
    explicit Cache(const TransCacheEntry& legendre): legendre_(legendre) {}
    const TransCacheEntry& legendre() const { return legendre_; }
private:
    TransCacheEntry legendre_;
};
class TransFactory {
public:
    TransFactory(const std::string& name, const std::string& backend): name_(name) { registry()[name] = this; (void)backend; }
    virtual ~TransFactory() = default;
    virtual const TransImpl* make(const Cache&, const Grid&, const Domain&, int, const eckit::Configuration&) = 0;
    static const TransImpl* build(const std::string& type, const Grid& g, int truncation, const eckit::Configuration& c) {
        return build(type, Cache(), g, truncation, c);
    }
    static const TransImpl* build(const std::string& type, const Cache& cache, const Grid& g, int truncation,
                                  const eckit::Configuration& c) {
        auto it = registry().find(type);
        if (it == registry().end()) throw eckit::Exception("no such Trans backend: " + type);
        return it->second->make(cache, g, g.domain(), truncation, c);
    }
    static bool has(const std::string& type) { return registry().count(type) != 0; }
private:
    static std::map<std::string, TransFactory*>& registry() {
        static std::map<std::string, TransFactory*> r;
        return r;
    }
    std::string name_;
};
template <class T>
class TransBuilderGrid : public TransFactory {
    const TransImpl* make(const Cache& cache, const Grid& grid, const Domain& domain, int truncation,
                          const eckit::Configuration& config) override {
        return new T(cache, grid, domain, truncation, config);
    }
public:
    TransBuilderGrid(const std::string& name, const std::string& backend): TransFactory(name, backend) {}
};
}  // namespace trans
}  // namespace atlas
