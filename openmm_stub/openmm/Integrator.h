#ifndef OPENMM_STUB_INTEGRATOR_H_
#define OPENMM_STUB_INTEGRATOR_H_
namespace OpenMM {

class Context;

class Integrator {
public:
    Integrator() : stepSize(0.001), owner(nullptr) {}
    virtual ~Integrator() {}
    double getStepSize() const { return stepSize; }
    void setStepSize(double size) { stepSize = size; }
    virtual void step(int steps) = 0;
protected:
    friend class Context;
    double stepSize;
    Context* owner;
};

} // namespace OpenMM
#endif
