// Mock of atlas/runtime/Exception.h (runtime/Exception.h:23-76): same entry points, std::runtime_error underneath.
#pragma once
#include <stdexcept>
#include <string>
namespace eckit {
struct CodeLocation {
    const char* file;
    int line;
};
struct Exception : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct NotImplemented : Exception {
    using Exception::Exception;
};
struct AssertionFailed : Exception {
    using Exception::Exception;
};
}  // namespace eckit
#define Here() ::eckit::CodeLocation{__FILE__, __LINE__}
namespace atlas {
[[noreturn]] inline void throw_NotImplemented(const std::string& m, const eckit::CodeLocation&) { throw eckit::NotImplemented(m); }
[[noreturn]] inline void throw_NotImplemented(const eckit::CodeLocation&) { throw eckit::NotImplemented("not implemented"); }
[[noreturn]] inline void throw_Exception(const std::string& m, const eckit::CodeLocation&) { throw eckit::Exception(m); }
}  // namespace atlas
#define ATLAS_NOTIMPLEMENTED ::atlas::throw_NotImplemented(Here())
#define ATLAS_ASSERT(cond, ...) \
    do {                        \
        if (!(cond)) throw ::eckit::AssertionFailed(#cond); \
    } while (0)
