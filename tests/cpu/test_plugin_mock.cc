// Compiles the plugin translation unit (plugin/atlas-b200/src/B200Plugin.cc) against the mock atlas headers and checks
// that its static initialisers register the plugin and the "b200" Trans backend, as loading the real plugin would.
#include <cstdio>

#include "../../plugin/atlas-b200/src/B200Plugin.cc"

int main() {
    const bool plugin = !atlas::Plugin::loaded().empty() && atlas::Plugin::loaded().front() == "atlas-b200";
    const bool backend = atlas::trans::TransFactory::has("b200");
    std::printf("plugin registered: %d, backend registered: %d\n", plugin, backend);
    return plugin && backend ? 0 : 1;
}
