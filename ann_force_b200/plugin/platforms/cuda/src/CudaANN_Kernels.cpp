// CudaCalcANN_ForceKernel over the C-ABI launcher.  Replaces
// /root/reference/platforms/cuda/src/CudaANN_Kernels.cpp:74-395 (upload + code generation + empty stubs).
#include "CudaANN_Kernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include <map>
#include <string>
#include <vector>

using namespace OpenMM;
using std::string;
using std::vector;

namespace {

// Keeps atoms of the selection together when OpenMM sorts molecules (role of CudaANN_ForceInfo,
// CudaANN_Kernels.cpp:46-72 — which left getParticlesInGroup empty).
class CudaANN_ForceInfo : public CudaForceInfo {
public:
    explicit CudaANN_ForceInfo(const ANN_Force& force) {
        for (int a : force.get_index_of_backbone_atoms()) particles.push_back(a - 1);
        for (const vector<int>& quad : force.get_list_of_index_of_atoms_forming_dihedrals()) particles.insert(particles.end(), quad.begin(), quad.end());
        for (const vector<int>& pair : force.get_list_of_pair_index_for_distances()) particles.insert(particles.end(), pair.begin(), pair.end());
    }
    int getNumParticleGroups() { return 1; }
    void getParticlesInGroup(int, vector<int>& out) { out = particles; }
    bool areGroupsIdentical(int, int) { return true; }
private:
    vector<int> particles;
};

class ReorderListener : public CudaContext::ReorderListener {
public:
    explicit ReorderListener(CudaCalcANN_ForceKernel& kernel) : kernel(kernel) {}
    void execute() { kernel.atomsWereReordered(); }
private:
    CudaCalcANN_ForceKernel& kernel;
};

void check(annf_status status, const char* what) {
    if (status != ANNF_OK) throw OpenMMException(string(what) + ": " + annf_last_error());
}

int layer_code(const string& type) {
    static const std::map<string, int> codes = {{"Linear", ANNF_LINEAR}, {"Tanh", ANNF_TANH}, {"Circular", ANNF_CIRCULAR}, {"Softmax", ANNF_SOFTMAX}};
    auto it = codes.find(type);
    if (it == codes.end()) throw OpenMMException("ANN_Force: unknown layer type '" + type + "'");   // the legacy map silently inserted 0
    return it->second;
}

} // namespace

CudaCalcANN_ForceKernel::~CudaCalcANN_ForceKernel() {
    cu.setAsCurrent();
    if (handle != nullptr) annf_destroy(handle);
}

void CudaCalcANN_ForceKernel::initialize(const System& system, const ANN_Force& force) {
    cu.setAsCurrent();
    force.validate(system.getNumParticles());
    const vector<int>& nodes = force.get_num_of_nodes();
    const int L = (int) nodes.size();
    vector<int> types(L - 1);
    vector<const double*> coeff(L - 1), bias(L - 1);
    for (int l = 0; l + 1 < L; l++) {
        types[l] = layer_code(force.get_layer_types()[l]);
        coeff[l] = force.get_coeffients_of_connections()[l].data();
        bias[l] = force.get_values_of_biased_nodes()[l].data();
    }
    vector<int> dihedrals, pairs;
    for (const vector<int>& quad : force.get_list_of_index_of_atoms_forming_dihedrals()) dihedrals.insert(dihedrals.end(), quad.begin(), quad.end());
    for (const vector<int>& pair : force.get_list_of_pair_index_for_distances()) pairs.insert(pairs.end(), pair.begin(), pair.end());

    annf_desc desc = {};
    desc.struct_size = sizeof(annf_desc);
    desc.device = cu.getDeviceIndex();
    desc.num_layers = L;
    desc.num_of_nodes = nodes.data();
    desc.layer_types = types.data();
    desc.coeff = coeff.data();
    desc.bias = bias.data();
    desc.potential_center = force.get_potential_center().data();
    desc.num_center = (int) force.get_potential_center().size();
    desc.force_constant = force.get_force_constant();
    desc.scaling_factor = force.get_scaling_factor();
    desc.data_type_in_input_layer = force.get_data_type_in_input_layer();
    desc.index_of_backbone_atoms = force.get_index_of_backbone_atoms().data();
    desc.num_backbone_atoms = (int) force.get_index_of_backbone_atoms().size();
    desc.dihedrals = dihedrals.data();
    desc.num_dihedrals = (int) dihedrals.size() / 4;
    desc.pairs = pairs.data();
    desc.num_pairs = (int) pairs.size() / 2;
    desc.num_system_atoms = system.getNumParticles();
    desc.precision = ANNF_PREC_AUTO;
    check(annf_create(&desc, &handle), "ANN_Force CUDA kernel");
    atomsWereReordered();
    cu.addReorderListener(new ReorderListener(*this));
    cu.addForce(new CudaANN_ForceInfo(force));
}

void CudaCalcANN_ForceKernel::atomsWereReordered() {
    if (handle == nullptr) return;
    check(annf_set_atom_order(handle, reinterpret_cast<const int*>(cu.getAtomIndexArray().getDevicePointer()),
                              system.getNumParticles(), cu.getCurrentStream()),
          "ANN_Force CUDA kernel (atom order)");
}

double CudaCalcANN_ForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    cu.setAsCurrent();
    unsigned flags = 0;
    if (cu.getUseDoublePrecision()) flags |= ANNF_FLAG_POSQ_DOUBLE;
    if (cu.getUseDoublePrecision() || cu.getUseMixedPrecision()) flags |= ANNF_FLAG_ENERGY_DOUBLE;
    if (!includeForces) flags |= ANNF_FLAG_SKIP_FORCES;
    if (!includeEnergy) flags |= ANNF_FLAG_SKIP_ENERGY;
    const void* correction = cu.getUseMixedPrecision() ? reinterpret_cast<const void*>(cu.getPosqCorrection().getDevicePointer()) : nullptr;
    check(annf_eval_openmm(handle, reinterpret_cast<const void*>(cu.getPosq().getDevicePointer()), correction,
                           reinterpret_cast<unsigned long long*>(cu.getForce().getDevicePointer()), cu.getPaddedNumAtoms(),
                           reinterpret_cast<void*>(cu.getEnergyBuffer().getDevicePointer()), flags, cu.getCurrentStream()),
          "ANN_Force CUDA kernel (execute)");
    return 0.0;   // like every OpenMM CUDA force: the energy lives in the device energy buffer
}

void CudaCalcANN_ForceKernel::copyParametersToContext(ContextImpl& context, const ANN_Force& force) {
    cu.setAsCurrent();
    check(annf_update_parameters(handle, force.get_potential_center().data(), (int) force.get_potential_center().size(),
                                 force.get_force_constant(), cu.getCurrentStream()),
          "updateParametersInContext");
}
