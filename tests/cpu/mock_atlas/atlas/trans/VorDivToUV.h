// Mock of trans/VorDivToUV.h:34-107: the abstract VorDivToUVImpl, the name-keyed factory and its builder template.
#pragma once
#include <map>
#include <string>
#include "atlas/functionspace/Spectral.h"
#include "atlas/trans/detail/TransImpl.h"
namespace atlas {
namespace trans {
class VorDivToUVImpl : public util::Object {
public:
    virtual ~VorDivToUVImpl() = default;
    virtual int truncation() const = 0;
    virtual void execute(const int nb_coeff, const int nb_fields, const double vorticity[], const double divergence[],
                         double U[], double V[], const eckit::Configuration& = util::NoConfig()) const = 0;
};
class VorDivToUVFactory {
public:
    static std::map<std::string, VorDivToUVFactory*>& registry() {
        static std::map<std::string, VorDivToUVFactory*> r;
        return r;
    }
    static bool has(const std::string& name) { return registry().count(name) != 0; }
    static VorDivToUVImpl* build(const std::string& name, int truncation, const eckit::Configuration& c = util::NoConfig()) {
        return registry().at(name)->make(truncation, c);
    }
    virtual VorDivToUVImpl* make(const FunctionSpace& sp, const eckit::Configuration&) = 0;
    virtual VorDivToUVImpl* make(int truncation, const eckit::Configuration&) = 0;
protected:
    explicit VorDivToUVFactory(const std::string& name) { registry()[name] = this; }
    virtual ~VorDivToUVFactory() = default;
};
template <class T>
class VorDivToUVBuilder : public VorDivToUVFactory {
    VorDivToUVImpl* make(const FunctionSpace& sp, const eckit::Configuration& config) override { return new T(sp, config); }
    VorDivToUVImpl* make(int truncation, const eckit::Configuration& config) override { return new T(truncation, config); }
public:
    explicit VorDivToUVBuilder(const std::string& name): VorDivToUVFactory(name) {}
};
}  // namespace trans
}  // namespace atlas
