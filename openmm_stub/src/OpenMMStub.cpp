// The few out-of-line members of the stand-in SDK (in real OpenMM they live in libOpenMM).
#include "openmm/Context.h"

namespace OpenMM {

ForceImpl& Force::getImplInContext(Context& context) {
    for (ForceImpl* impl : context.impl->getForceImpls())
        if (&impl->getOwner() == this) return *impl;
    throw OpenMMException("getImplInContext: This Force is not present in the Context");
}

ContextImpl& Force::getContextImpl(Context& context) { return *context.impl; }

} // namespace OpenMM
