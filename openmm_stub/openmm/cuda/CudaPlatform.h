#ifndef OPENMM_STUB_CUDAPLATFORM_H_
#define OPENMM_STUB_CUDAPLATFORM_H_
#include "openmm/Platform.h"
#include "openmm/cuda/CudaContext.h"
#include "openmm/internal/ContextImpl.h"
#include <map>
#include <string>
#include <vector>

namespace OpenMM {

// CUDA platform of the stand-in SDK.  Context properties: "Precision" = single (default) | mixed | double,
// "DeviceIndex", and the stub-only "StubReverseAtomOrder" = true to start from a non-trivial atom order.
class CudaPlatform : public Platform, public StubPlatformHooks {
public:
    class PlatformData {
    public:
        std::vector<CudaContext*> contexts;
        ~PlatformData() { for (CudaContext* c : contexts) delete c; }
    };
    const std::string& getName() const { static const std::string name = "CUDA"; return name; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {
        auto get = [&](const std::string& key, const std::string& fallback) {
            auto it = properties.find(key);
            return it == properties.end() ? fallback : it->second;
        };
        PlatformData* data = new PlatformData();
        data->contexts.push_back(new CudaContext(context.getSystem(), std::stoi(get("DeviceIndex", "0")), get("Precision", "single"),
                                                 get("StubReverseAtomOrder", "false") == "true"));
        context.setPlatformData(data);
    }
    void contextDestroyed(ContextImpl& context) const { delete static_cast<PlatformData*>(context.getPlatformData()); }
    void stubSetPositions(ContextImpl& context, const std::vector<Vec3>& positions) const { cu(context).setPositions(positions); }
    void stubGetPositions(ContextImpl& context, std::vector<Vec3>& positions) const { cu(context).getPositions(positions); }
    void stubGetForces(ContextImpl& context, std::vector<Vec3>& forces) const { cu(context).getForces(forces); }
    void stubBeginComputation(ContextImpl& context) const { cu(context).clearBuffers(); }
    double stubFinishComputation(ContextImpl& context) const { return cu(context).sumEnergy(); }
    static CudaContext& cu(ContextImpl& context) { return *static_cast<PlatformData*>(context.getPlatformData())->contexts[0]; }
};

} // namespace OpenMM
#endif
