#ifndef OPENMM_STUB_REALTYPE_H_
#define OPENMM_STUB_REALTYPE_H_
typedef double RealOpenMM;
#endif
