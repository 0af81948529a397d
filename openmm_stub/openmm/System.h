#ifndef OPENMM_STUB_SYSTEM_H_
#define OPENMM_STUB_SYSTEM_H_
#include "openmm/Force.h"
#include <vector>

namespace OpenMM {

class System {
public:
    System() {}
    ~System() { for (Force* f : forces) delete f; }
    int getNumParticles() const { return (int) masses.size(); }
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size() - 1; }
    double getParticleMass(int index) const { return masses.at(index); }
    int addForce(Force* force) { forces.push_back(force); return (int) forces.size() - 1; }   // takes ownership
    int getNumForces() const { return (int) forces.size(); }
    Force& getForce(int index) { return *forces.at(index); }
    const Force& getForce(int index) const { return *forces.at(index); }
private:
    System(const System&);
    System& operator=(const System&);
    std::vector<double> masses;
    std::vector<Force*> forces;
};

} // namespace OpenMM
#endif
