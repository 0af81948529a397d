#ifndef OPENMM_ANN_H_
#define OPENMM_ANN_H_
// Umbrella include of the ANN_Force plugin API (same name as the reference's openmmapi/include/OpenMM_ANN.h).
#include "openmm/ANN_Force.h"
#include "openmm/ANN_Kernels.h"
#include "openmm/internal/ANN_ForceImpl.h"
#endif
