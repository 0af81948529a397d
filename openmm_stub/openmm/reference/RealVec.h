#ifndef OPENMM_STUB_REALVEC_H_
#define OPENMM_STUB_REALVEC_H_
#include "openmm/Vec3.h"
namespace OpenMM { typedef Vec3 RealVec; }
#endif
