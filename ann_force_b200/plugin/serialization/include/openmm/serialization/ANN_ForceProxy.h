#ifndef OPENMM_ANN_FORCE_PROXY_H_
#define OPENMM_ANN_FORCE_PROXY_H_
// Serialization proxy for ANN_Force, so that a System containing the bias force can be written to and read from XML
// (XmlSerializer) like any built-in force.  The reference has none — its README lists it as a TODO
// (/root/reference/README.md:74) — which is SURVEY.md section 8f rank 4.
#include "openmm/internal/windowsExport.h"
#include "openmm/serialization/SerializationProxy.h"

namespace OpenMM {

class OPENMM_EXPORT ANN_ForceProxy : public SerializationProxy {
public:
    ANN_ForceProxy();
    void serialize(const void* object, SerializationNode& node) const;
    void* deserialize(const SerializationNode& node) const;
};

} // namespace OpenMM

/** Registers the proxy; also run by a static initialiser when the API library is loaded. */
extern "C" OPENMM_EXPORT void registerANNSerializationProxies();

#endif
