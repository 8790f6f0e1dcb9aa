// ham_host.h -- the host half of rimu_ham_create: validation of a rimu_ham_desc and construction of the HamDev image
// (scalars, constant tables, neighbour table).  Pure host code, shared by api.cu (which uploads the tables) and by the
// host-emulation test build (tests/cuda/host_ham.cpp), so that what the kernels receive is what the CPU tests check.
#pragma once
#include "../../include/rimu_b200.h"
#include "hamiltonians.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

struct HamHostImage {
    HamDev dev;                       // pointers are left null: the caller points them at device or host copies of the tables
    int hk = -1, W = 1;
    std::vector<double> tables;       // kes | ws | us | pot  (offsets 0, T, 2T, 3T with T = RIMU_MAX_TABLE_MODES)
    std::vector<unsigned char> nbr;   // HubbardRealSpace: nbr[(site-1)*nnb + dir] = neighbour site (1-based) or 0
    std::string error;
};

static inline int ham_neighbor_site(const rimu_ham_desc *d, int mode, int chosen) { // geometry.jl:161-175,232-235
    int D = d->ndim, idx = mode - 1, x[3];
    for (int k = 0; k < D; k++) { x[k] = idx % d->dims[k] + 1; idx /= d->dims[k]; }
    if (chosen <= D) x[chosen - 1] += 1; else x[chosen - D - 1] -= 1;
    for (int k = 0; k < D; k++) {
        if (d->fold[k]) { x[k] = ((x[k] - 1) % d->dims[k] + d->dims[k]) % d->dims[k] + 1; }
        else if (x[k] < 1 || x[k] > d->dims[k]) return 0;
    }
    int lin = 0, stride = 1;
    for (int k = 0; k < D; k++) { lin += (x[k] - 1) * stride; stride *= d->dims[k]; }
    return lin + 1;
}

static inline int ham_host_fail(HamHostImage *img, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    img->error = buf;
    return RIMU_ERR_INVALID;
}

// returns 0 or RIMU_ERR_INVALID (message in img->error)
static inline int ham_build_host(const rimu_ham_desc *d, HamHostImage *img) {
    const int M = d->num_modes, kind = d->addr_kind, model = d->model;
    if (M < 1 || M > RIMU_MAX_MODES) return ham_host_fail(img, "num_modes %d unsupported (1..%d)", M, RIMU_MAX_MODES);
    int hk = -1, bits = 0;
    if (kind == RIMU_ADDR_BOSE) {
        if (d->num_components != 1) return ham_host_fail(img, "BoseFS must have one component");
        bits = d->num_particles[0] + M - 1;
        if (bits + 1 > 128) return ham_host_fail(img, "BoseFS{%d,%d} needs %d bits; at most 127 supported", d->num_particles[0], M, bits);
        if (model == RIMU_HUBBARD_REAL_1D || model == RIMU_HUBBARD_REAL_1D_EP || model == RIMU_EXTENDED_HUBBARD_REAL_1D) hk = HK_REAL1D_BOSE;
        else if (model == RIMU_HUBBARD_MOM_1D || model == RIMU_EXTENDED_HUBBARD_MOM_1D || model == RIMU_HUBBARD_MOM_1D_EP) hk = HK_MOM1D_BOSE;
        else if (model == RIMU_HUBBARD_REAL_SPACE) hk = HK_RS_BOSE;
    } else if (kind == RIMU_ADDR_FERMI) {
        if (d->num_components != 1) return ham_host_fail(img, "FermiFS must have one component");
        bits = M;
        if (M > 63) return ham_host_fail(img, "FermiFS with more than 63 modes unsupported");
        if (model == RIMU_HUBBARD_REAL_SPACE) hk = HK_RS_FERMI;
    } else if (kind == RIMU_ADDR_FERMI2C) {
        if (d->num_components != 2) return ham_host_fail(img, "FermiFS2C must have two components");
        bits = 2 * M;
        if (M > 32) return ham_host_fail(img, "two-component fermions with more than 32 modes unsupported");
        if (d->num_particles[0] == M && d->num_particles[1] == M && M == 32)
            return ham_host_fail(img, "completely filled 32-mode two-component address collides with the empty-slot sentinel");
        if (model == RIMU_HUBBARD_MOM_1D || model == RIMU_HUBBARD_MOM_1D_EP) hk = HK_MOM1D_F2C;
        else if (model == RIMU_HUBBARD_REAL_SPACE) hk = HK_RS_F2C;
        else if (model == RIMU_TRANSCORRELATED_1D) hk = HK_TC_F2C;
    }
    else if (kind == RIMU_ADDR_COMPOSITE) { // general CompositeFS: components packed side by side from the low bits
        const int C = d->num_components;
        if (C < 2 || C > RIMU_MAX_COMPONENTS) return ham_host_fail(img, "CompositeFS must have 2..%d components", RIMU_MAX_COMPONENTS);
        for (int c = 0; c < C; c++) {
            const int n = d->comp_particles[c];
            if (d->comp_kind[c] == RIMU_ADDR_BOSE) { if (n < 0) return ham_host_fail(img, "negative particle number"); bits += n + M - 1; }
            else if (d->comp_kind[c] == RIMU_ADDR_FERMI) {
                if (n < 0 || n > M) return ham_host_fail(img, "FermiFS component with %d particles in %d modes", n, M);
                if (M > 64) return ham_host_fail(img, "FermiFS components with more than 64 modes unsupported");
                bits += M;
            } else return ham_host_fail(img, "components of a CompositeFS must be BoseFS or FermiFS");
        }
        if (bits + 1 > 128) return ham_host_fail(img, "CompositeFS needs %d bits; at most 127 supported", bits);
        if (model == RIMU_HUBBARD_REAL_SPACE) hk = HK_RS_COMP;
    }
    if (hk < 0)
        return ham_host_fail(img, "model %d is not implemented for address kind %d (no CPU fallback exists)", model, kind);
    if ((hk == HK_MOM1D_BOSE || hk == HK_MOM1D_F2C || hk == HK_TC_F2C) && M > RIMU_MAX_TABLE_MODES)
        return ham_host_fail(img, "momentum-space models support at most %d modes", RIMU_MAX_TABLE_MODES);
    if (hk == HK_MOM1D_BOSE && M < 3) return ham_host_fail(img, "HubbardMom1D needs at least 3 modes");
    { // the device decoders index off-diagonals with 32-bit arithmetic
        double n1 = d->num_particles[0], n2 = d->num_particles[1], m = M, lmax = 0;
        if (hk == HK_MOM1D_BOSE) lmax = n1 * (n1 - 1) * (m - 2) + 2 * n1 * (m - 1);
        else if (hk == HK_TC_F2C) lmax = n1 * n2 * (m - 1) + (n1 * (n1 - 1) * n2 + n2 * (n2 - 1) * n1) * m * m;
        else if (hk == HK_MOM1D_F2C) lmax = n1 * n2 * (m - 1) + (n1 + n2) * (m - 1);
        else if (hk == HK_RS_COMP) { lmax = 0; for (int c = 0; c < d->num_components; c++) lmax += 6.0 * d->comp_particles[c]; }
        else lmax = (n1 + n2) * 6;
        if (lmax >= 2147483648.0) return ham_host_fail(img, "more than 2^31 off-diagonals per address are unsupported");
    }
    img->hk = hk;
    img->W = (kind == RIMU_ADDR_BOSE || kind == RIMU_ADDR_COMPOSITE) ? ((bits + 1 + 63) / 64) : 1;
    HamDev &v = img->dev;
    memset(&v, 0, sizeof(v));
    v.hk = hk; v.M = M; v.N0 = d->num_particles[0]; v.N1 = d->num_particles[1];
    v.ndim = d->ndim; v.nnb = 2 * d->ndim; v.cutoff = d->cutoff; v.three_body = d->three_body_term; v.has_pot = d->has_potential;
    v.u = d->u; v.t = d->t; v.v = d->v; v.tc0 = d->t_comp[0]; v.tc1 = d->t_comp[1];
    v.u00 = d->u_mat[0]; v.u10 = d->u_mat[1];
    v.u_2m = d->u / (2 * M); v.u_m = d->u / M;
    v.variant = model == RIMU_HUBBARD_REAL_1D_EP ? 1 : model == RIMU_EXTENDED_HUBBARD_REAL_1D ? 2
              : model == RIMU_EXTENDED_HUBBARD_MOM_1D ? 1 : model == RIMU_HUBBARD_MOM_1D_EP ? 2 : 0;
    v.v_m = d->v / M;
    v.bc = d->boundary_condition;
    if (v.variant == 0) { // the plain models get the kernels compiled with the variant folded in
        if (hk == HK_REAL1D_BOSE) hk = HK_REAL1D_BOSE_PLAIN;
        else if (hk == HK_MOM1D_BOSE) hk = HK_MOM1D_BOSE_PLAIN;
        img->hk = hk; v.hk = hk;
    }
    if (v.variant == 2 && (v.bc < 0 || v.bc > 2)) return ham_host_fail(img, "invalid boundary condition");
    int nz = 0;
    if (hk == HK_RS_COMP) {
        const int C = d->num_components;
        v.ncomp = C;
        int off = 0;
        for (int c = 0; c < C; c++) {
            v.cbose[c] = d->comp_kind[c] == RIMU_ADDR_BOSE;
            v.cbits[c] = v.cbose[c] ? d->comp_particles[c] + M - 1 : M;
            v.coff[c] = off; off += v.cbits[c];
            v.tcs[c] = d->comp_t[c];
            for (int c2 = 0; c2 < C; c2++) {
                v.umat[c + C * c2] = d->comp_u[c + C * c2];
                nz += d->comp_u[c + C * c2] != 0.0;
                if (d->comp_u[c + C * c2] != d->comp_u[c2 + C * c]) return ham_host_fail(img, "`u` must be symmetric");
            }
        }
    } else
    for (int i = 0; i < d->num_components * d->num_components; i++) nz += d->u_mat[(i % d->num_components) + 2 * (i / d->num_components)] != 0.0;
    v.umat_zero = nz == 0;
    // tables: kes | ws | us | pot
    img->tables.assign(3 * RIMU_MAX_TABLE_MODES + 2 * RIMU_MAX_MODES, 0.0);
    memcpy(&img->tables[0], d->kes, sizeof(d->kes));
    memcpy(&img->tables[RIMU_MAX_TABLE_MODES], d->ws, sizeof(d->ws));
    memcpy(&img->tables[2 * RIMU_MAX_TABLE_MODES], d->us, sizeof(d->us));
    memcpy(&img->tables[3 * RIMU_MAX_TABLE_MODES], d->potential, sizeof(d->potential));
    img->nbr.clear();
    if (model == RIMU_HUBBARD_REAL_SPACE) {
        if (d->ndim < 1 || d->ndim > 3) return ham_host_fail(img, "geometry must have 1..3 dimensions");
        int prod = 1;
        for (int k = 0; k < d->ndim; k++) prod *= d->dims[k];
        if (prod != M) return ham_host_fail(img, "`geometry` does not have the correct number of sites");
        img->nbr.resize((size_t)M * v.nnb);
        for (int s = 1; s <= M; s++)
            for (int c = 1; c <= v.nnb; c++) img->nbr[(size_t)(s - 1) * v.nnb + (c - 1)] = (unsigned char)ham_neighbor_site(d, s, c);
    }
    return 0;
}

// point the table pointers of a HamDev at (device or host) copies of img's tables
static inline void ham_set_tables(HamDev *v, const double *tables, const unsigned char *nbr) {
    v->kes = tables; v->ws = tables + RIMU_MAX_TABLE_MODES; v->us = tables + 2 * RIMU_MAX_TABLE_MODES;
    v->pot = tables + 3 * RIMU_MAX_TABLE_MODES;
    v->nbr = nbr;
}
