#ifndef OPENMM_STUB_ASSERTIONUTILITIES_H_
#define OPENMM_STUB_ASSERTIONUTILITIES_H_
#include "openmm/OpenMMException.h"
#include <cmath>
#include <sstream>
#include <string>

namespace OpenMM {
inline void throwException(const char* file, int line, const std::string& details) {
    std::stringstream message;
    message << "Assertion failure at " << file << ":" << line;
    if (details.size() > 0) message << ".  " << details;
    throw OpenMMException(message.str());
}
} // namespace OpenMM

#define ASSERT(cond) {if (!(cond)) OpenMM::throwException(__FILE__, __LINE__, "");};
#define ASSERT_EQUAL(expected, found) {if (!((expected) == (found))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_TOL(expected, found, tol) {double _scale_ = std::abs(expected) > 1.0 ? std::abs(expected) : 1.0; if (!(std::abs((expected)-(found))/_scale_ <= (tol))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_VEC(expected, found, tol) {ASSERT_EQUAL_TOL((expected)[0], (found)[0], (tol)); ASSERT_EQUAL_TOL((expected)[1], (found)[1], (tol)); ASSERT_EQUAL_TOL((expected)[2], (found)[2], (tol));};
#endif
