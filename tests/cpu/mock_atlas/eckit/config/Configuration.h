// Minimal stand-in for eckit::Configuration (only what TransB200.h touches). Test infrastructure.
#pragma once
#include <map>
#include <string>
namespace eckit {
class Configuration {
public:
    virtual ~Configuration() = default;
    bool get(const std::string& key, int& v) const {
        auto it = ints_.find(key);
        if (it == ints_.end()) return false;
        v = it->second;
        return true;
    }
    bool get(const std::string& key, std::string& v) const {
        auto it = strings_.find(key);
        if (it == strings_.end()) return false;
        v = it->second;
        return true;
    }
    bool getBool(const std::string& key, bool dflt) const {
        int v = 0;
        return get(key, v) ? v != 0 : dflt;
    }
    Configuration& set(const std::string& key, int v) {
        ints_[key] = v;
        return *this;
    }
    Configuration& set(const std::string& key, const std::string& v) {
        strings_[key] = v;
        return *this;
    }
private:
    std::map<std::string, int> ints_;
    std::map<std::string, std::string> strings_;
};
}  // namespace eckit
