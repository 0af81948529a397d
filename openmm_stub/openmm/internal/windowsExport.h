#ifndef OPENMM_STUB_WINDOWSEXPORT_H_
#define OPENMM_STUB_WINDOWSEXPORT_H_
#ifndef OPENMM_EXPORT
#define OPENMM_EXPORT __attribute__((visibility("default")))
#endif
#endif
