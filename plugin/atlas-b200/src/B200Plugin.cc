// atlas plugin that makes `option::type("b200")` available to any atlas program without touching libatlas:
// atlas loads the shared libraries it finds through ATLAS_PLUGIN_PATH / ATLAS_PLUGINS (ecmwf/atlas
// src/atlas/library/Library.cc:172-174), loading this one runs the two static initialisers below -- the plugin object
// (library/Plugin.h:20-27) and the Trans builder (trans/detail/TransFactory.h:114-129, the idiom of
// trans/local/TransLocal.cc:57) -- and from then on
//
//     atlas::trans::Trans trans(grid, truncation, atlas::option::type("b200"));
//
// constructs an atlas::trans::TransB200 (include/atlas_b200/TransB200.h), i.e. a plan of libsptrans_b200.so.
#include <string>

#include "atlas/library/Plugin.h"
#include "atlas_b200/LegendreCacheCreatorB200.h"
#include "atlas_b200/TransB200.h"
#include "atlas_b200/VorDivToUVB200.h"

namespace atlas_b200_plugin {

class B200Plugin : public atlas::Plugin {
public:
    B200Plugin(): atlas::Plugin("atlas-b200") {}
    static const B200Plugin& instance() {
        static B200Plugin plugin;
        return plugin;
    }
    std::string version() const override { return "0.1.0"; }
    std::string gitsha1(unsigned int) const override { return "not available"; }
};

REGISTER_LIBRARY(B200Plugin);

namespace {
// backend name "b200", registered for Trans(grid, truncation, config) like "local" and "ectrans" are
atlas::trans::TransBuilderGrid<atlas::trans::TransB200> register_trans_b200("b200", "b200");
// ... and for Trans(gp_functionspace, sp_functionspace, config): the key atlas looks up is
// type + "(" + gp.type() + "," + sp.type() + ")" (trans/detail/TransFactory.cc:206-211), as TransIFSStructuredColumns.cc:35-39
atlas::trans::TransBuilderFunctionSpace<atlas::trans::TransB200> register_trans_b200_fs("b200(StructuredColumns,Spectral)", "b200");
// LegendreCacheCreator(grid, truncation, option::type("b200")): same idiom as trans/local/LegendreCacheCreatorLocal.cc:30
atlas::trans::LegendreCacheCreatorBuilder<atlas::trans::LegendreCacheCreatorB200> register_cache_creator_b200("b200");
// VorDivToUV(truncation, option::type("b200")): same idiom as trans/local/VorDivToUVLocal.cc:25
atlas::trans::VorDivToUVBuilder<atlas::trans::VorDivToUVB200> register_vordiv_to_uv_b200("b200");
}  // namespace

}  // namespace atlas_b200_plugin
