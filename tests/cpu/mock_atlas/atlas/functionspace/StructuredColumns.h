// Mock of functionspace::StructuredColumns (functionspace/StructuredColumns.h): grid() and sizeOwned()
// (functionspace/detail/StructuredColumns.h:104-105,137).
#pragma once
#include "atlas/functionspace/Spectral.h"
namespace atlas {
namespace functionspace {
class StructuredColumns {
public:
    explicit StructuredColumns(const FunctionSpace& fs): fs_(fs) {}
    explicit operator bool() const { return fs_.type() == "StructuredColumns"; }
    const Grid& grid() const { return fs_.grid(); }
    idx_t sizeOwned() const { return fs_.owned(); }
    static std::string type() { return "StructuredColumns"; }
private:
    FunctionSpace fs_;
};
}  // namespace functionspace
}  // namespace atlas
