#ifndef OPENMM_STUB_PLATFORM_H_
#define OPENMM_STUB_PLATFORM_H_
#include "openmm/Kernel.h"
#include "openmm/KernelFactory.h"
#include "openmm/OpenMMException.h"
#include <map>
#include <string>
#include <vector>

namespace OpenMM {

class ContextImpl;

// A compute back end: a registry of kernel factories keyed by kernel name.
class Platform {
public:
    virtual ~Platform() {
        // factories may be shared between names; delete each once
        std::vector<KernelFactory*> seen;
        for (auto& kv : factories) {
            bool dup = false;
            for (auto* f : seen) dup = dup || (f == kv.second);
            if (!dup) { seen.push_back(kv.second); delete kv.second; }
        }
    }
    virtual const std::string& getName() const = 0;
    virtual double getSpeed() const { return 1.0; }
    virtual bool supportsDoublePrecision() const { return true; }
    // hooks the Context calls so that a platform can attach / free its per-context data
    virtual void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {}
    virtual void contextDestroyed(ContextImpl& context) const {}

    void registerKernelFactory(const std::string& name, KernelFactory* factory) { factories[name] = factory; }
    bool supportsKernels(const std::vector<std::string>& names) const {
        for (auto& n : names) if (factories.find(n) == factories.end()) return false;
        return true;
    }
    Kernel createKernel(const std::string& name, ContextImpl& context) const {
        auto it = factories.find(name);
        if (it == factories.end())
            throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel: " + name);
        return Kernel(it->second->createKernelImpl(name, *this, context));
    }

    static void registerPlatform(Platform* platform) { registry().push_back(platform); }
    static int getNumPlatforms() { return (int) registry().size(); }
    static Platform& getPlatform(int index) { return *registry().at(index); }
    static Platform& getPlatformByName(const std::string& name) {
        for (auto* p : registry()) if (p->getName() == name) return *p;
        throw OpenMMException("There is no registered Platform called \"" + name + "\"");
    }
    // plugin loading is the host application's job; the stub links plugins statically
    static std::vector<std::string> loadPluginsFromDirectory(const std::string&) { return std::vector<std::string>(); }
private:
    static std::vector<Platform*>& registry() { static std::vector<Platform*> r; return r; }
    std::map<std::string, KernelFactory*> factories;
};

} // namespace OpenMM
#endif
