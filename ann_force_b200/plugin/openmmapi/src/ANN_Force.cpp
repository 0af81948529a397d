// ANN_Force: parameter container.  Semantics follow /root/reference/openmmapi/src/ANN_Force.cpp
// (1-based atom numbers in, dihedral / pair lists stored 0-based and appended to), sized from the arguments.
#include "openmm/ANN_Force.h"
#include "openmm/internal/ANN_ForceImpl.h"
#include <cmath>
#include <sstream>

using namespace OpenMM;
using std::string;
using std::vector;

ANN_Force::ANN_Force() : force_constant(0.0), scaling_factor(1.0), data_type_in_input_layer(0) {}
ANN_Force::~ANN_Force() {}

const vector<int>& ANN_Force::get_num_of_nodes() const { return num_of_nodes; }
void ANN_Force::set_num_of_nodes(vector<int> num) { num_of_nodes.swap(num); }

const vector<int>& ANN_Force::get_index_of_backbone_atoms() const { return index_of_backbone_atoms; }
void ANN_Force::set_index_of_backbone_atoms(vector<int> indices) { index_of_backbone_atoms.swap(indices); }

const vector<vector<double> >& ANN_Force::get_coeffients_of_connections() const { return coeff; }
void ANN_Force::set_coeffients_of_connections(vector<vector<double> > coefficients) { coeff.swap(coefficients); }

const vector<string>& ANN_Force::get_layer_types() const { return layer_types; }
void ANN_Force::set_layer_types(vector<string> temp_layer_types) { layer_types.swap(temp_layer_types); }

const vector<vector<double> >& ANN_Force::get_values_of_biased_nodes() const { return values_of_biased_nodes; }
void ANN_Force::set_values_of_biased_nodes(vector<vector<double> > bias) { values_of_biased_nodes.swap(bias); }

const vector<double>& ANN_Force::get_potential_center() const { return potential_center; }
void ANN_Force::set_potential_center(vector<double> temp_potential_center) { potential_center.swap(temp_potential_center); }

const double ANN_Force::get_force_constant() const { return force_constant; }
void ANN_Force::set_force_constant(double temp_force_constant) { force_constant = temp_force_constant; }

const double ANN_Force::get_scaling_factor() const { return scaling_factor; }
void ANN_Force::set_scaling_factor(double temp_scaling_factor) { scaling_factor = temp_scaling_factor; }

const int ANN_Force::get_data_type_in_input_layer() const { return data_type_in_input_layer; }
void ANN_Force::set_data_type_in_input_layer(int temp_data_type_in_input_layer) { data_type_in_input_layer = temp_data_type_in_input_layer; }

const vector<vector<int> >& ANN_Force::get_list_of_index_of_atoms_forming_dihedrals() const {
    return list_of_index_of_atoms_forming_dihedrals;
}

void ANN_Force::set_list_of_index_of_atoms_forming_dihedrals(vector<vector<int> > temp_list_of_index) {
    for (const vector<int>& quad : temp_list_of_index) {
        if (quad.size() < 4) throw OpenMMException("set_list_of_index_of_atoms_forming_dihedrals: a dihedral needs four atoms");
        vector<int> zero_based(4);
        for (int j = 0; j < 4; j++) zero_based[j] = quad[j] - 1;      // PDB numbering starts at 1
        list_of_index_of_atoms_forming_dihedrals.push_back(zero_based);
    }
}

// phi / psi of consecutive residues from a backbone given as N, CA, C triples: the same quadruples, in the
// same order, as /root/reference/openmmapi/src/ANN_Force.cpp:118-154
void ANN_Force::set_list_of_index_of_atoms_forming_dihedrals_from_index_of_backbone_atoms(vector<int> backbone) {
    const int residues = (int) backbone.size() / 3;
    for (int res = 0; res < residues; res++) {
        const int first_of_quad[2] = {3 * res - 1, 3 * res};              // phi starts at C(i-1), psi at N(i)
        const bool present[2] = {res != 0, res != residues - 1};
        for (int which = 0; which < 2; which++) {
            if (!present[which]) continue;
            vector<int> quad(4);
            for (int j = 0; j < 4; j++) quad[j] = backbone[first_of_quad[which] + j] - 1;
            list_of_index_of_atoms_forming_dihedrals.push_back(quad);
        }
    }
}

const vector<vector<int> >& ANN_Force::get_list_of_pair_index_for_distances() const { return list_of_pair_index_for_distances; }

void ANN_Force::set_list_of_pair_index_for_distances(vector<vector<int> > temp_list_of_index) {
    for (const vector<int>& pair : temp_list_of_index) {
        if (pair.size() < 2) throw OpenMMException("set_list_of_pair_index_for_distances: a pair needs two atoms");
        list_of_pair_index_for_distances.push_back(vector<int>{pair[0] - 1, pair[1] - 1});
    }
}

void ANN_Force::validate(int numParticles) const {
    std::stringstream msg;
    const int L = (int) num_of_nodes.size();
    if (L < 2) throw OpenMMException("ANN_Force: set_num_of_nodes needs at least an input and an output layer");
    if ((int) layer_types.size() != L - 1 || (int) coeff.size() != L - 1 || (int) values_of_biased_nodes.size() != L - 1) {
        msg << "ANN_Force: " << L << " layers need " << L - 1 << " layer types, coefficient sets and bias sets (got "
            << layer_types.size() << ", " << coeff.size() << ", " << values_of_biased_nodes.size() << ")";
        throw OpenMMException(msg.str());
    }
    for (int l = 0; l + 1 < L; l++) {
        const string& t = layer_types[l];
        if (t != "Linear" && t != "Tanh" && t != "Circular" && t != "Softmax")
            throw OpenMMException("ANN_Force: unknown layer type '" + t + "'");
        if ((long long) coeff[l].size() != (long long) num_of_nodes[l] * num_of_nodes[l + 1]) {
            msg << "ANN_Force: connection " << l << " needs " << (long long) num_of_nodes[l] * num_of_nodes[l + 1]
                << " coefficients, got " << coeff[l].size();
            throw OpenMMException(msg.str());
        }
        if ((int) values_of_biased_nodes[l].size() != num_of_nodes[l + 1])
            throw OpenMMException("ANN_Force: bias vector size does not match the layer size");
        if (t == "Circular" && num_of_nodes[l + 1] % 2 != 0)
            throw OpenMMException("ANN_Force: a Circular layer needs an even number of nodes");
    }
    const int last = num_of_nodes[L - 1];
    const int want = layer_types[L - 2] == "Circular" ? last / 2 : last;
    if ((int) potential_center.size() != want) {
        msg << "ANN_Force: potential_center needs " << want << " values, got " << potential_center.size();
        throw OpenMMException(msg.str());
    }
    if (scaling_factor == 0.0 || !std::isfinite(scaling_factor)) throw OpenMMException("ANN_Force: scaling_factor must be finite and non-zero");
    if (data_type_in_input_layer == 1) {
        if (3 * (int) index_of_backbone_atoms.size() != num_of_nodes[0])
            throw OpenMMException("ANN_Force: Cartesian input needs 3 * len(index_of_backbone_atoms) == num_of_nodes[0]");
        std::vector<char> seen((size_t) numParticles + 1, 0);
        for (int a : index_of_backbone_atoms) {
            if (a < 1 || a > numParticles) throw OpenMMException("ANN_Force: index_of_backbone_atoms is 1-based and must lie inside the System");
            // the reference accumulates per list entry (ANN_ReferenceKernels.cpp:401); the batched kernels write one force row
            // per selected atom, so a repeated atom is rejected instead of being evaluated two different ways
            if (seen[a]) throw OpenMMException("ANN_Force: index_of_backbone_atoms lists an atom more than once");
            seen[a] = 1;
        }
    } else if (data_type_in_input_layer == 0) {
        if (2 * (int) list_of_index_of_atoms_forming_dihedrals.size() != num_of_nodes[0])
            throw OpenMMException("ANN_Force: dihedral input needs 2 * number of dihedrals == num_of_nodes[0]");
    } else if (data_type_in_input_layer == 2) {
        if ((int) list_of_pair_index_for_distances.size() != num_of_nodes[0])
            throw OpenMMException("ANN_Force: pair-distance input needs number of pairs == num_of_nodes[0]");
    } else {
        throw OpenMMException("ANN_Force: data_type_in_input_layer must be 0, 1 or 2");
    }
}

ForceImpl* ANN_Force::createImpl() const { return new ANN_ForceImpl(*this); }

void ANN_Force::updateParametersInContext(Context& context) {
    dynamic_cast<ANN_ForceImpl&>(getImplInContext(context)).updateParametersInContext(getContextImpl(context));
}
