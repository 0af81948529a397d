/* TEST INFRASTRUCTURE — CPU restatement ("port") of the ANN_Force hot path.  Not shipped, not timed
 * as the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load it.
 *
 * Plain C99, fp64 throughout (the reference's RealOpenMM is double).  Each function cites the lines
 * of /root/reference it re-states.  Parity of this file is PINNED (tests/test_oracle.py) against
 *   - the reference's own known-answer vectors (openmmapi/tests/test_ANN_package.cpp:121,156,92-93),
 *   - golden outputs of the reference's unmodified kernel (tests/golden/, made by
 *     tests/golden/make_golden.py from oracle/_ref), and
 *   - oracle/_ref itself on random networks whenever that library is present.
 *
 * Layer type codes: 0 Linear, 1 Tanh, 2 Circular, 3 Softmax (the reference compares strings,
 * ANN_ReferenceKernels.cpp:200-238).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int num_layers;               /* node layers including the input layer (reference NUM_OF_LAYERS) */
    const int* num_of_nodes;      /* [num_layers] */
    const int* layer_types;       /* [num_layers-1] */
    const double* const* coeff;   /* [num_layers-1], row-major [n_out][n_in]  (ANN_ReferenceKernels.cpp:81-83) */
    const double* const* bias;    /* [num_layers-1] */
    const double* potential_center;
    int num_center;
    double force_constant;
    double scaling_factor;
    int data_type_in_input_layer; /* 0 dihedral cos/sin, 1 Cartesian, 2 pair distances */
    const int* index_of_backbone_atoms; /* 1-based (ANN_ReferenceKernels.cpp:132) */
    int num_backbone;
    const int* dihedrals;         /* 0-based quadruples, i.e. as stored after the setter's -1 (ANN_Force.cpp:111) */
    int num_dihedrals;
    const int* pairs;             /* 0-based pairs (ANN_Force.cpp:164) */
    int num_pairs;
} ann_oracle_net;

enum { LIN = 0, TANH = 1, CIRC = 2, SOFTMAX = 3 };

typedef struct {
    int L;
    double** z; /* "input_of_each_layer"  : pre-activations  */
    double** a; /* "output_of_each_layer" : activations      */
    double** d; /* "derivatives_of_each_layer" : dE/d a      */
} workspace;

static workspace ws_alloc(const ann_oracle_net* net) {
    workspace w;
    w.L = net->num_layers;
    w.z = (double**) calloc(w.L, sizeof(double*));
    w.a = (double**) calloc(w.L, sizeof(double*));
    w.d = (double**) calloc(w.L, sizeof(double*));
    for (int l = 0; l < w.L; l++) {
        w.z[l] = (double*) calloc(net->num_of_nodes[l], sizeof(double));
        w.a[l] = (double*) calloc(net->num_of_nodes[l], sizeof(double));
        w.d[l] = (double*) calloc(net->num_of_nodes[l], sizeof(double));
    }
    return w;
}

static void ws_free(workspace* w) {
    for (int l = 0; l < w->L; l++) { free(w->z[l]); free(w->a[l]); free(w->d[l]); }
    free(w->z); free(w->a); free(w->d);
}

/* ---- vec3 helpers.  vdiv multiplies by the reciprocal like OpenMM's Vec3::operator/ ---- */
static void vsub(const double* a, const double* b, double* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static double vdot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void vcross(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static void vdiv(const double* a, double s, double* o) { double r = 1.0 / s; o[0] = a[0] * r; o[1] = a[1] * r; o[2] = a[2] * r; }

/* Dihedral sign convention of the reference (ANN_ReferenceKernels.cpp:456,612): the sign of
 * (sum of components of n1 x n2) * (sum of components of the middle bond vector). */
static int dihedral_sign(const double* n1_cross_n2, const double* d2) {
    return (n1_cross_n2[0] + n1_cross_n2[1] + n1_cross_n2[2]) * (d2[0] + d2[1] + d2[2]) > 0 ? 1 : -1;
}

/* cos / sin of one dihedral: ANN_ReferenceKernels.cpp:599-636 */
void ann_oracle_cos_sin_four_atoms(const double* p0, const double* p1, const double* p2, const double* p3,
                                   double* cos_value, double* sin_value) {
    double d1[3], d2[3], d3[3], n1[3], n2[3], sv[3];
    vsub(p1, p0, d1); vsub(p2, p1, d2); vsub(p3, p2, d3);
    vcross(d1, d2, n1); vcross(d2, d3, n2);
    vdiv(n1, sqrt(vdot(n1, n1)), n1);
    vdiv(n2, sqrt(vdot(n2, n2)), n2);
    *cos_value = vdot(n1, n2);
    vcross(n1, n2, sv);
    *sin_value = sqrt(vdot(sv, sv)) * dihedral_sign(sv, d2);
}

/* Feature stage of candidate_2: ANN_ReferenceKernels.cpp:124-160 */
static void features(const ann_oracle_net* net, const double* pos, double* x) {
    const double s = net->scaling_factor;
    if (net->data_type_in_input_layer == 0) {                    /* :125-127, :579-597 */
        for (int q = 0; q < net->num_dihedrals; q++) {
            const int* id = net->dihedrals + 4 * q;
            ann_oracle_cos_sin_four_atoms(pos + 3 * id[0], pos + 3 * id[1], pos + 3 * id[2], pos + 3 * id[3],
                                          &x[2 * q], &x[2 * q + 1]);
        }
    } else if (net->data_type_in_input_layer == 1) {             /* :128-151 */
        const int n = 3 * net->num_backbone;
        double com[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < net->num_backbone; i++)
            for (int c = 0; c < 3; c++)
                x[3 * i + c] = pos[3 * (net->index_of_backbone_atoms[i] - 1) + c];
        for (int i = 0; i < n; i++) {                            /* :138-141, same operation order */
            x[i] /= s;
            com[i % 3] += x[i] / n * 3;
        }
        for (int i = 0; i < n; i++) x[i] -= com[i % 3];          /* :143-145 */
    } else {                                                     /* :152-160 */
        for (int p = 0; p < net->num_pairs; p++) {
            double a[3], b[3], d[3];
            vdiv(pos + 3 * net->pairs[2 * p], s, a);
            vdiv(pos + 3 * net->pairs[2 * p + 1], s, b);
            vsub(a, b, d);
            x[p] = sqrt(vdot(d, d));
        }
    }
}

/* calculate_output_of_each_layer: ANN_ReferenceKernels.cpp:179-260 */
static void forward(const ann_oracle_net* net, workspace* w, const double* x) {
    memcpy(w->a[0], x, sizeof(double) * net->num_of_nodes[0]);
    for (int l = 1; l < net->num_layers; l++) {
        const int n_in = net->num_of_nodes[l - 1], n_out = net->num_of_nodes[l];
        const double* W = net->coeff[l - 1];
        for (int j = 0; j < n_out; j++) {                        /* :193-198 */
            double acc = net->bias[l - 1][j];
            for (int i = 0; i < n_in; i++) acc += W[(size_t) j * n_in + i] * w->a[l - 1][i];
            w->z[l][j] = acc;
        }
        switch (net->layer_types[l - 1]) {
        case LIN:                                                /* :200-204 */
            for (int j = 0; j < n_out; j++) w->a[l][j] = w->z[l][j];
            break;
        case TANH:                                               /* :205-209 */
            for (int j = 0; j < n_out; j++) w->a[l][j] = tanh(w->z[l][j]);
            break;
        case CIRC:                                               /* :210-225 */
            for (int j = 0; j < n_out / 2; j++) {
                double p = w->z[l][2 * j], q = w->z[l][2 * j + 1];
                double r = sqrt(p * p + q * q);
                w->a[l][2 * j] = p / r;
                w->a[l][2 * j + 1] = q / r;
            }
            break;
        default: {                                               /* Softmax, no max-shift: :226-234 */
            double sum = 0.0;
            for (int j = 0; j < n_out; j++) sum += exp(w->z[l][j]);
            for (int j = 0; j < n_out; j++) w->a[l][j] = exp(w->z[l][j]) / sum;
        }
        }
    }
}

/* Signed angle of a (cos, sin) output pair and its periodic distance to the centre.
 * ANN_ReferenceKernels.h:100-107 and ANN_ReferenceKernels.cpp:277-291; 6.2832 is the reference's 2*pi. */
static double circular_angle(double cos_value, double sin_value, int* sign) {
    *sign = sin_value > 0 ? 1 : -1;
    return acos(cos_value) * (*sign);
}

/* update_and_get_potential_energy: ANN_ReferenceKernels.h:89-119 */
static double energy(const ann_oracle_net* net, const workspace* w, const double* center) {
    const int L = net->num_layers, n = net->num_of_nodes[L - 1];
    const double k = net->force_constant;
    double e = 0.0;
    if (net->layer_types[L - 2] != CIRC) {                       /* :92-97 */
        for (int j = 0; j < n; j++) e += 0.5 * k * (w->a[L - 1][j] - center[j]) * (w->a[L - 1][j] - center[j]);
    } else {                                                     /* :98-116 */
        for (int j = 0; j < n / 2; j++) {
            int sign;
            double angle = circular_angle(w->a[L - 1][2 * j], w->a[L - 1][2 * j + 1], &sign);
            double d1 = angle - center[j], d2 = d1 + 6.2832, d3 = d1 - 6.2832;
            double m = fmin(fmin(d1 * d1, d2 * d2), d3 * d3);
            e += 0.5 * k * m;
        }
    }
    return e;
}

/* back_prop: ANN_ReferenceKernels.cpp:262-387 */
static void backward(const ann_oracle_net* net, workspace* w, const double* center) {
    const int L = net->num_layers, n_last = net->num_of_nodes[L - 1];
    const double k = net->force_constant;
    for (int l = 0; l < L; l++) memcpy(w->d[l], w->a[l], sizeof(double) * net->num_of_nodes[l]);  /* :263 */
    if (net->layer_types[L - 2] != CIRC) {                       /* :265-270 */
        for (int j = 0; j < n_last; j++) w->d[L - 1][j] = (w->a[L - 1][j] - center[j]) * k;
    } else {                                                     /* :271-296 */
        for (int j = 0; j < n_last / 2; j++) {
            double c = w->a[L - 1][2 * j], s = w->a[L - 1][2 * j + 1];
            int sign;
            double angle = circular_angle(c, s, &sign);
            double d1 = angle - center[j], d2 = d1 + 6.2832, d3 = d1 - 6.2832;
            double dist = d1;
            if (fabs(d2) < fabs(dist)) dist = d2;
            if (fabs(d3) < fabs(dist)) dist = d3;
            w->d[L - 1][2 * j] = k * dist * (-1) / sqrt(1 - c * c) * sign;   /* dE/dcos :292-293 */
            w->d[L - 1][2 * j + 1] = 0;                                       /* dE/dsin :294 */
        }
    }
    for (int l = L - 2; l >= 0; l--) {                           /* :299-373 */
        const int n_in = net->num_of_nodes[l], n_out = net->num_of_nodes[l + 1];
        const double* W = net->coeff[l];
        double* dz = (double*) malloc(sizeof(double) * n_out);   /* dE/d(pre-activation) of layer l+1 */
        switch (net->layer_types[l]) {
        case CIRC:                                               /* :307-318 */
            for (int j = 0; j < n_out / 2; j++) {
                double p = w->z[l + 1][2 * j], q = w->z[l + 1][2 * j + 1];
                double r = sqrt(p * p + q * q);
                double dp = w->d[l + 1][2 * j], dq = w->d[l + 1][2 * j + 1];
                dz[2 * j] = q / (r * r * r) * (q * dp - p * dq);
                dz[2 * j + 1] = p / (r * r * r) * (p * dq - q * dp);
            }
            break;
        case TANH:                                               /* :343-348 */
            for (int j = 0; j < n_out; j++) dz[j] = w->d[l + 1][j] * (1 - w->a[l + 1][j] * w->a[l + 1][j]);
            break;
        case LIN:                                                /* :349-354 */
            for (int j = 0; j < n_out; j++) dz[j] = w->d[l + 1][j];
            break;
        default:                                                 /* Softmax Jacobian :355-365 */
            for (int j = 0; j < n_out; j++) {
                double t = 0.0;
                for (int s = 0; s < n_out; s++) {
                    t += -w->a[l + 1][j] * w->a[l + 1][s] * w->d[l + 1][s];
                    if (j == s) t += w->a[l + 1][s] * w->d[l + 1][s];
                }
                dz[j] = t;
            }
        }
        for (int m = 0; m < n_in; m++) {                         /* W^T * dz  (:326-336, :340-371) */
            double acc = 0.0;
            for (int j = 0; j < n_out; j++) acc += W[(size_t) j * n_in + m] * dz[j];
            w->d[l][m] = acc;
        }
        free(dz);
    }
}

/* Chain rule from d(cos), d(sin) of one dihedral to its four atoms.
 * Re-derives, in vector form, what ANN_ReferenceKernels.cpp:425-577 spells out component by
 * component (the sympy-generated block :466-477 is d{cos,sin}/d{n1,n2}; :480-534 is d{n1,n2}/d{bond
 * vectors}; :546-562 scatters to atoms):
 *   cos = n1.n2/(|n1||n2|),  sin = sign*|n1 x n2|/(|n1||n2|),  n1 = d1 x d2,  n2 = d2 x d3. */
static void dihedral_forces(const double* pos, const int* id, double g_cos, double g_sin, double* frc) {
    double d1[3], d2[3], d3[3], n1[3], n2[3], c[3];
    vsub(pos + 3 * id[1], pos + 3 * id[0], d1);
    vsub(pos + 3 * id[2], pos + 3 * id[1], d2);
    vsub(pos + 3 * id[3], pos + 3 * id[2], d3);
    vcross(d1, d2, n1); vcross(d2, d3, n2); vcross(n1, n2, c);
    const int sign = dihedral_sign(c, d2);
    const double a = vdot(n1, n1), b = vdot(n2, n2), n1n2 = vdot(n1, n2), cc = vdot(c, c);
    const double la = sqrt(a), lb = sqrt(b), lc = sqrt(cc);
    double n2xc[3], cxn1[3], g1[3], g2[3];
    vcross(n2, c, n2xc); vcross(c, n1, cxn1);
    for (int i = 0; i < 3; i++) {
        double dcos_dn1 = n2[i] / (la * lb) - n1[i] * n1n2 / (a * la * lb);
        double dcos_dn2 = n1[i] / (la * lb) - n2[i] * n1n2 / (b * la * lb);
        double dsin_dn1 = sign * (n2xc[i] / (lc * la * lb) - n1[i] * lc / (a * la * lb));
        double dsin_dn2 = sign * (cxn1[i] / (lc * la * lb) - n2[i] * lc / (b * la * lb));
        g1[i] = g_cos * dcos_dn1 + g_sin * dsin_dn1;   /* dE/dn1 */
        g2[i] = g_cos * dcos_dn2 + g_sin * dsin_dn2;   /* dE/dn2 */
    }
    double e1[3], e2a[3], e2b[3], e3[3];               /* dE/d(bond vectors) */
    vcross(d2, g1, e1);                                /* n1 = d1 x d2  ->  dE/dd1 = d2 x g1 */
    vcross(g1, d1, e2a); vcross(d3, g2, e2b);          /* dE/dd2 = g1 x d1 + d3 x g2 */
    vcross(g2, d2, e3);                                /* dE/dd3 = g2 x d2 */
    for (int i = 0; i < 3; i++) {                      /* :546-553: force[idx[k]] += t, force[idx[k+1]] -= t */
        double e2 = e2a[i] + e2b[i];
        frc[3 * id[0] + i] += e1[i];
        frc[3 * id[1] + i] += e2 - e1[i];
        frc[3 * id[2] + i] += e3[i] - e2;
        frc[3 * id[3] + i] += -e3[i];
    }
}

/* get_all_forces_from_derivative_of_first_layer: ANN_ReferenceKernels.cpp:389-423.  Accumulates (+=). */
static void scatter_forces(const ann_oracle_net* net, const double* pos, const double* g, double* frc) {
    const double s = net->scaling_factor;
    if (net->data_type_in_input_layer == 0) {                    /* :392-396 */
        for (int q = 0; q < net->num_of_nodes[0] / 2; q++)
            dihedral_forces(pos, net->dihedrals + 4 * q, g[2 * q], g[2 * q + 1], frc);
    } else if (net->data_type_in_input_layer == 1) {             /* :397-410 */
        const int A = net->num_backbone;
        for (int i = 0; i < A; i++)
            for (int c = 0; c < 3; c++) {
                double* f = &frc[3 * (net->index_of_backbone_atoms[i] - 1) + c];
                *f += -g[3 * i + c] / s;
                for (int kk = 0; kk < A; kk++) *f += g[3 * kk + c] / s / A;   /* Jacobian of centroid removal, same order as :403-406 */
            }
    } else {                                                     /* :411-421 */
        for (int p = 0; p < net->num_pairs; p++) {
            const int i0 = net->pairs[2 * p], i1 = net->pairs[2 * p + 1];
            double d[3];
            vsub(pos + 3 * i0, pos + 3 * i1, d);
            double dist = sqrt(vdot(d, d));
            double f = g[p] / (s * dist);
            for (int c = 0; c < 3; c++) { frc[3 * i0 + c] -= d[c] * f; frc[3 * i1 + c] += d[c] * f; }
        }
    }
}

/* R independent evaluations.  positions [R][num_atoms][3]; centers [R][num_center] or NULL;
 * forces [R][num_atoms][3] OVERWRITTEN with this force's contribution; energies [R].
 * Follows candidate_2 (ANN_ReferenceKernels.cpp:122-171): features, forward, back_prop, forces, energy. */
int ann_oracle_eval_batch(const ann_oracle_net* net, int num_atoms, int R, const double* positions,
                          const double* centers, double* forces, double* energies) {
    workspace w = ws_alloc(net);
    double* x = (double*) calloc(net->num_of_nodes[0] > 0 ? net->num_of_nodes[0] : 1, sizeof(double));
    for (int r = 0; r < R; r++) {
        const double* pos = positions + (size_t) r * num_atoms * 3;
        double* frc = forces + (size_t) r * num_atoms * 3;
        const double* center = centers ? centers + (size_t) r * net->num_center : net->potential_center;
        memset(frc, 0, sizeof(double) * num_atoms * 3);
        features(net, pos, x);
        forward(net, &w, x);
        backward(net, &w, center);
        scatter_forces(net, pos, w.d[0], frc);
        energies[r] = energy(net, &w, center);
    }
    free(x);
    ws_free(&w);
    return 0;
}

/* Unit-level entry used by the known-answer tests (test_ANN_package.cpp:98-173):
 * forward over an explicit input vector, then back_prop; both concatenated over layers 0..L-1. */
int ann_oracle_forward_backward(const ann_oracle_net* net, const double* input, double* layer_outputs,
                                double* layer_derivs) {
    workspace w = ws_alloc(net);
    forward(net, &w, input);
    backward(net, &w, net->potential_center);
    size_t o = 0;
    for (int l = 0; l < net->num_layers; l++)
        for (int j = 0; j < net->num_of_nodes[l]; j++, o++) {
            layer_outputs[o] = w.a[l][j];
            layer_derivs[o] = w.d[l][j];
        }
    ws_free(&w);
    return 0;
}
