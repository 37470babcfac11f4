// Mock of atlas::Plugin (library/Plugin.h:20-27) and of eckit's REGISTER_LIBRARY: enough to compile the plugin
// translation unit plugin/atlas-b200/src/B200Plugin.cc and to observe that loading it registers the plugin.
#pragma once
#include <string>
#include <vector>
namespace atlas {
class Plugin {
public:
    explicit Plugin(const std::string& name): name_(name) {}
    virtual ~Plugin() = default;
    const std::string& name() const { return name_; }
    virtual std::string version() const = 0;
    virtual std::string gitsha1(unsigned int count) const = 0;
    virtual void init() {}
    static std::vector<std::string>& loaded() {
        static std::vector<std::string> v;
        return v;
    }
private:
    std::string name_;
};
}  // namespace atlas
#define REGISTER_LIBRARY(X)                                        \
    static const bool X##_registered = [] {                        \
        ::atlas::Plugin::loaded().push_back(X::instance().name()); \
        return true;                                               \
    }()
