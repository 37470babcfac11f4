// Mock of trans/LegendreCacheCreator.h:34-118: the abstract creator, its name-keyed factory and the builder template.
#pragma once
#include <map>
#include <string>
#include "atlas/grid.h"
#include "atlas/trans/Cache.h"
#include "atlas/trans/detail/TransImpl.h"
namespace atlas {
namespace trans {
class LegendreCacheCreatorImpl : public util::Object {
public:
    virtual ~LegendreCacheCreatorImpl() = default;
    virtual bool supported() const = 0;
    virtual std::string uid() const = 0;
    virtual void create(const std::string& path) const = 0;
    virtual Cache create() const = 0;
    virtual size_t estimate() const = 0;
};
class LegendreCacheCreatorFactory {
public:
    static std::map<std::string, LegendreCacheCreatorFactory*>& registry() {
        static std::map<std::string, LegendreCacheCreatorFactory*> r;
        return r;
    }
    static bool has(const std::string& name) { return registry().count(name) != 0; }
    static LegendreCacheCreatorImpl* build(const std::string& name, const Grid& g, int truncation,
                                           const eckit::Configuration& c = util::NoConfig()) {
        return registry().at(name)->make(g, truncation, c);
    }
    virtual LegendreCacheCreatorImpl* make(const Grid&, int truncation, const eckit::Configuration&) = 0;
protected:
    explicit LegendreCacheCreatorFactory(const std::string& name) { registry()[name] = this; }
    virtual ~LegendreCacheCreatorFactory() = default;
};
template <class T>
class LegendreCacheCreatorBuilder : public LegendreCacheCreatorFactory {
    LegendreCacheCreatorImpl* make(const Grid& grid, int truncation, const eckit::Configuration& config) override {
        return new T(grid, truncation, config);
    }
public:
    explicit LegendreCacheCreatorBuilder(const std::string& name): LegendreCacheCreatorFactory(name) {}
};
}  // namespace trans
}  // namespace atlas
