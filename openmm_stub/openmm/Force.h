#ifndef OPENMM_STUB_FORCE_H_
#define OPENMM_STUB_FORCE_H_
#include "openmm/internal/windowsExport.h"
#include <map>
#include <string>
#include <vector>

namespace OpenMM {

class Context;
class ContextImpl;
class ForceImpl;

// Base class of every force term.  A System owns the Force objects added to it; a Context asks
// each one for a ForceImpl at construction time.
class Force {
public:
    Force() : forceGroup(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int group) { forceGroup = group; }
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
protected:
    friend class ContextImpl;
    virtual ForceImpl* createImpl() const = 0;
    ForceImpl& getImplInContext(Context& context);
    ContextImpl& getContextImpl(Context& context);
private:
    int forceGroup;
};

} // namespace OpenMM
#endif
