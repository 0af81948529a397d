#ifndef OPENMM_STUB_OPENMM_H_
#define OPENMM_STUB_OPENMM_H_
#include "openmm/Vec3.h"
#include "openmm/OpenMMException.h"
#include "openmm/Force.h"
#include "openmm/System.h"
#include "openmm/Platform.h"
#include "openmm/Context.h"
#include "openmm/State.h"
#include "openmm/VerletIntegrator.h"
#endif
