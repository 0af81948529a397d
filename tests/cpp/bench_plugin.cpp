// Per-step cost of the bias force THROUGH THE C++ PLUGIN LAYER (ANN_ForceImpl::calcForcesAndEnergy ->
// CudaCalcANN_ForceKernel::execute -> annf_eval_openmm), i.e. what OpenMM pays per MD step.  BASELINE config 2:
// alanine dipeptide, 22 atoms, 7 backbone heavy atoms selected, 21-40-4 Tanh/Circular, mixed precision.
// Prints one JSON object.  usage: bench_plugin [steps]
#include "OpenMM.h"
#include "OpenMM_ANN.h"
#include "CudaANN_KernelFactory.h"
#include "openmm/cuda/CudaPlatform.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

using namespace OpenMM;
using namespace std;

static double lcg(unsigned long long& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return ((s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0; }

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 20000;
    registerANN_CudaKernelFactories();
    System system;
    const int num_atoms = 22;
    for (int i = 0; i < num_atoms; i++) system.addParticle(12.0);
    ANN_Force* f = new ANN_Force();
    const vector<int> nodes = {21, 40, 4};
    f->set_num_of_nodes(nodes);
    f->set_layer_types({"Tanh", "Circular"});
    unsigned long long seed = 1234;
    vector<vector<double> > coeff(2), bias(2);
    for (int l = 0; l < 2; l++) {
        const double bound = sqrt(6.0 / (nodes[l] + nodes[l + 1]));
        for (int i = 0; i < nodes[l] * nodes[l + 1]; i++) coeff[l].push_back(bound * lcg(seed));
        for (int i = 0; i < nodes[l + 1]; i++) bias[l].push_back(0.1 * lcg(seed));
    }
    f->set_coeffients_of_connections(coeff);
    f->set_values_of_biased_nodes(bias);
    f->set_potential_center({0.7, -1.1});
    f->set_force_constant(500);
    f->set_scaling_factor(0.5);
    f->set_index_of_backbone_atoms({2, 5, 7, 9, 15, 17, 19});     // ACE-ALA-NME backbone heavy atoms
    f->set_data_type_in_input_layer(1);
    system.addForce(f);
    VerletIntegrator integrator(0.002);
    map<string, string> props; props["Precision"] = "mixed";
    Context context(system, integrator, Platform::getPlatformByName("CUDA"), props);
    vector<Vec3> positions(num_atoms);
    for (int i = 0; i < num_atoms; i++) positions[i] = Vec3(0.5 * lcg(seed), 0.5 * lcg(seed), 0.5 * lcg(seed));
    context.setPositions(positions);
    ContextImpl& impl = context.getImpl();
    ForceImpl* force_impl = impl.getForceImpls()[0];
    CudaContext& cu = CudaPlatform::cu(impl);
    const double e_check = context.getState(State::Energy).getPotentialEnergy();
    for (int i = 0; i < 200; i++) force_impl->calcForcesAndEnergy(impl, true, true, 0xFFFFFFFF);
    cudaStreamSynchronize(cu.getCurrentStream());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto t0 = chrono::steady_clock::now();
    cudaEventRecord(e0, cu.getCurrentStream());
    for (int i = 0; i < steps; i++) force_impl->calcForcesAndEnergy(impl, true, true, 0xFFFFFFFF);
    cudaEventRecord(e1, cu.getCurrentStream());
    cudaStreamSynchronize(cu.getCurrentStream());
    auto t1 = chrono::steady_clock::now();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double wall_us = chrono::duration<double, micro>(t1 - t0).count() / steps;
    printf("{\"steps\": %d, \"gpu_us_per_step\": %.4f, \"wall_us_per_step\": %.4f, \"energy\": %.12g, "
           "\"path\": \"ANN_ForceImpl::calcForcesAndEnergy -> CudaCalcANN_ForceKernel::execute -> annf_eval_openmm\"}\n",
           steps, 1e3 * ms / steps, wall_us, e_check);
    return 0;
}
