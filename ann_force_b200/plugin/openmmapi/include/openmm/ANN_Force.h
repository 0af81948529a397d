#ifndef OPENMM_ANN_FORCE_H_
#define OPENMM_ANN_FORCE_H_
// ANN_Force: the user-facing parameter object of the autoencoder bias force.
//
// Source-compatible with the reference class (/root/reference/openmmapi/include/openmm/ANN_Force.h:51-138):
// every setter / getter keeps its name (including "coeffients"), signature and meaning, so existing C++ and
// SWIG/Python callers compile unchanged.  What differs is inside:
//   * the network depth is whatever set_num_of_nodes() is given — the reference fixes it with the
//     NUM_OF_LAYERS macro and its setters truncate or overflow for any other depth (ANN_Force.cpp:13-18,46-64).
//     NUM_OF_LAYERS / NUM_OF_BACKBONE_ATOMS stay defined only so that old code naming them still compiles;
//   * validate() reports inconsistent parameters with an OpenMMException instead of assert()
//     (ANN_ReferenceKernels.cpp:61-66,74) or a printf (:235-238).
#include "openmm/Force.h"
#include "openmm/OpenMMException.h"
#include "openmm/Vec3.h"
#include <string>
#include <vector>

#ifndef NUM_OF_LAYERS
#define NUM_OF_LAYERS 3            // legacy default depth; NOT a limit here
#endif
#ifndef NUM_OF_BACKBONE_ATOMS
#define NUM_OF_BACKBONE_ATOMS 6    // legacy default; NOT a limit here
#endif

namespace OpenMM {

class OPENMM_EXPORT ANN_Force : public Force {
public:
    ANN_Force();
    ~ANN_Force();

    const std::vector<int>& get_num_of_nodes() const;
    void set_num_of_nodes(std::vector<int> num);

    const std::vector<int>& get_index_of_backbone_atoms() const;               // 1-based, as given
    void set_index_of_backbone_atoms(std::vector<int> indices);

    const std::vector<std::vector<double> >& get_coeffients_of_connections() const;   // row-major [n_out][n_in] per connection
    void set_coeffients_of_connections(std::vector<std::vector<double> > coefficients);

    const std::vector<std::string>& get_layer_types() const;                   // "Linear" | "Tanh" | "Circular" | "Softmax"
    void set_layer_types(std::vector<std::string> temp_layer_types);

    const std::vector<std::vector<double> >& get_values_of_biased_nodes() const;
    void set_values_of_biased_nodes(std::vector<std::vector<double> > bias);

    const std::vector<double>& get_potential_center() const;
    void set_potential_center(std::vector<double> temp_potential_center);

    const double get_force_constant() const;
    void set_force_constant(double temp_force_constant);

    const double get_scaling_factor() const;
    void set_scaling_factor(double temp_scaling_factor);

    const int get_data_type_in_input_layer() const;                            // 0 cos/sin of dihedrals, 1 Cartesian, 2 pair distances
    void set_data_type_in_input_layer(int temp_data_type_in_input_layer);

    const std::vector<std::vector<int> >& get_list_of_index_of_atoms_forming_dihedrals() const;   // stored 0-based
    void set_list_of_index_of_atoms_forming_dihedrals(std::vector<std::vector<int> > temp_list_of_index);   // takes 1-based, appends
    void set_list_of_index_of_atoms_forming_dihedrals_from_index_of_backbone_atoms(std::vector<int> index_of_backbone_atoms);

    const std::vector<std::vector<int> >& get_list_of_pair_index_for_distances() const;           // stored 0-based
    void set_list_of_pair_index_for_distances(std::vector<std::vector<int> > temp_list_of_index);    // takes 1-based, appends

    /** Push the umbrella centre and force constant of this object into an existing Context
     *  (an empty stub in the reference on both platforms; implemented here). */
    void updateParametersInContext(Context& context);

    bool usesPeriodicBoundaryConditions() const { return false; }

    /** Throws OpenMMException describing the first inconsistency; called by every kernel's initialize(). */
    void validate(int numParticles) const;

protected:
    ForceImpl* createImpl() const;

private:
    std::vector<int> num_of_nodes;
    std::vector<int> index_of_backbone_atoms;
    std::vector<std::vector<int> > list_of_index_of_atoms_forming_dihedrals;
    std::vector<std::vector<int> > list_of_pair_index_for_distances;
    std::vector<std::vector<double> > coeff;
    std::vector<std::string> layer_types;
    std::vector<std::vector<double> > values_of_biased_nodes;
    std::vector<double> potential_center;
    double force_constant;
    double scaling_factor;
    int data_type_in_input_layer;
};

} // namespace OpenMM
#endif
