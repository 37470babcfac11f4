// Mock of atlas/util/Config.h: util::Config is the concrete, settable eckit::Configuration (eckit::LocalConfiguration).
#pragma once
#include "eckit/config/Configuration.h"
namespace atlas {
namespace util {
class Config : public eckit::Configuration {};
}  // namespace util
}  // namespace atlas
