// Registers the ANN_Force serialization proxy when the API library is loaded (the convention of OpenMM plugins).
#include "openmm/serialization/ANN_ForceProxy.h"
#include "openmm/ANN_Force.h"

using namespace OpenMM;

extern "C" OPENMM_EXPORT void registerANNSerializationProxies() {
    static ANN_ForceProxy proxy;
    SerializationProxy::registerProxy(typeid(ANN_Force), &proxy);
}

namespace {
struct RegisterAtLoad {
    RegisterAtLoad() { registerANNSerializationProxies(); }
} registerAtLoad;
}  // namespace
