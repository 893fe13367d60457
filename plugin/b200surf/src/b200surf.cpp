/*
 * b200surf.cpp -- plugin registration (pattern: plugins/doc_template/src/megamolplugin.cpp).
 */
#include "mmcore/factories/AbstractPluginInstance.h"
#include "mmcore/factories/PluginRegister.h"

#include "IsoSurfaceB200.h"
#include "ParticlesToDensityB200.h"

namespace megamol::b200surf {
class PluginInstance : public megamol::core::factories::AbstractPluginInstance {
    REGISTERPLUGIN(PluginInstance)
public:
    PluginInstance()
            : megamol::core::factories::AbstractPluginInstance(
                  "b200surf", "B200-native particle -> density volume -> isosurface path (libmmsurf, sm_100a)"){};

    ~PluginInstance() override = default;

    void registerClasses() override {
        this->module_descriptions.RegisterAutoDescription<megamol::b200surf::ParticlesToDensityB200>();
        this->module_descriptions.RegisterAutoDescription<megamol::b200surf::IsoSurfaceB200>();
    }
};
} // namespace megamol::b200surf
