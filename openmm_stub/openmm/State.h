#ifndef OPENMM_STUB_STATE_H_
#define OPENMM_STUB_STATE_H_
#include "openmm/Vec3.h"
#include <vector>

namespace OpenMM {

class State {
public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16 };
    State() : potentialEnergy(0.0), kineticEnergy(0.0) {}
    const std::vector<Vec3>& getPositions() const { return positions; }
    const std::vector<Vec3>& getVelocities() const { return velocities; }
    const std::vector<Vec3>& getForces() const { return forces; }
    double getPotentialEnergy() const { return potentialEnergy; }
    double getKineticEnergy() const { return kineticEnergy; }
private:
    friend class Context;
    std::vector<Vec3> positions, velocities, forces;
    double potentialEnergy, kineticEnergy;
};

} // namespace OpenMM
#endif
