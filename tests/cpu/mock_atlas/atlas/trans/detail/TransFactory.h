// Mock of the backend registry (trans/detail/TransFactory.h:114-129, util/Factory.h:94-113).
#pragma once
#include <map>
#include <string>
#include "atlas/functionspace/StructuredColumns.h"
#include "atlas/trans/Cache.h"
#include "atlas/trans/detail/TransImpl.h"
namespace atlas {
namespace trans {
class TransFactory {
public:
    TransFactory(const std::string& name, const std::string& backend): name_(name) { registry()[name] = this; (void)backend; }
    virtual ~TransFactory() = default;
    virtual const TransImpl* make(const Cache&, const Grid&, const Domain&, int, const eckit::Configuration&) = 0;
    virtual const TransImpl* make(const Cache&, const FunctionSpace&, const FunctionSpace&, const eckit::Configuration&) = 0;
    // Trans(gp, sp, config): key = type + "(" + gp.type() + "," + sp.type() + ")"  (trans/detail/TransFactory.cc:206-211)
    static const TransImpl* build(const std::string& type, const FunctionSpace& gp, const FunctionSpace& sp,
                                  const eckit::Configuration& c) {
        auto it = registry().find(type + "(" + gp.type() + "," + sp.type() + ")");
        if (it == registry().end()) throw eckit::Exception("no such Trans backend: " + type);
        return it->second->make(Cache(), gp, sp, c);
    }
    static const TransImpl* build(const std::string& type, const Grid& g, int truncation, const eckit::Configuration& c) {
        return build(type, Cache(), g, truncation, c);
    }
    static const TransImpl* build(const std::string& type, const Cache& cache, const Grid& g, int truncation,
                                  const eckit::Configuration& c) {
        auto it = registry().find(type);
        if (it == registry().end()) throw eckit::Exception("no such Trans backend: " + type);
        return it->second->make(cache, g, g.domain(), truncation, c);
    }
    // Trans(cache, grid, domain, truncation, config): trans/detail/TransFactory.cc:236-258
    static const TransImpl* build(const std::string& type, const Cache& cache, const Grid& g, int truncation,
                                  const eckit::Configuration& c, const Domain& domain) {
        auto it = registry().find(type);
        if (it == registry().end()) throw eckit::Exception("no such Trans backend: " + type);
        return it->second->make(cache, g, domain, truncation, c);
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
    const TransImpl* make(const Cache&, const FunctionSpace&, const FunctionSpace&, const eckit::Configuration&) override {
        throw eckit::Exception("This function should not be called");
    }
public:
    TransBuilderGrid(const std::string& name, const std::string& backend): TransFactory(name, backend) {}
};
template <class T>
class TransBuilderFunctionSpace : public TransFactory {  // trans/detail/TransFactory.h:99-113
    const TransImpl* make(const Cache& cache, const FunctionSpace& gp, const FunctionSpace& sp,
                          const eckit::Configuration& config) override {
        return new T(cache, gp, sp, config);
    }
    const TransImpl* make(const Cache&, const Grid&, const Domain&, int, const eckit::Configuration&) override {
        throw eckit::Exception("This function should not be called");
    }
public:
    TransBuilderFunctionSpace(const std::string& name, const std::string& backend): TransFactory(name, backend) {}
};
}  // namespace trans
}  // namespace atlas
