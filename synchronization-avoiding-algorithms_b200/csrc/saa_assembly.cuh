// saa_assembly.cuh — sparse, row-owned assembly of the per-rank stiffness on the device (kernel family K6).
//
// Restates /root/reference/Tools/Mat_construction.py:122-150 (Local_assembly_for_stiffness) together with
// Local_K_coronary (:79-119) WITHOUT the dense (3n)x(3n) accumulator: every local node owns its three matrix
// rows, walks its incident tetrahedra in ascending local element order (the order of the `K[P,Q] +=` loop,
// :125-148) and accumulates the 3x3 node blocks into its own slots — deterministic, no atomics.  Exact zeros
// are dropped and columns are ascending, like csr_matrix(dense) (:150).
//
// Element stiffness (linear tetrahedron, Shape_function_Deriv.py:33-67; Voigt order xx,yy,zz,yz,xz,xy,
// Mat_construction.py:99-104; D of commons.py:25-31):
//     (B_a^T D B_b)[A][B] = lmd*ga[A]*gb[B] + mu*ga[B]*gb[A] + mu*delta_AB*(ga . gb)
// times detJ*w, added once per quadrature point (four equal addends, w = 0.25/6, Qudrature.py:7-12).
// The reference evaluates B^T D B through BLAS; this closed form agrees with it to a few 1e-16 relative
// (not bit for bit) — see DESIGN.md "device assembly".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SAA_ASM_MAX_NB 128     // most distinct neighbour nodes (incl. itself) a node may have

struct SaaTetGeom {
    double g[4][3];   // physical gradients of the four shape functions
    double detJ;      // signed Jacobian determinant (= 6 * volume)
};

// Jacobian J[i][j] = P[j+1][i] - P[0][i] (Shape_function_Deriv.py:60-67), N_xyz = dN/dxi @ inv(J) (Mat_construction.py:96)
__device__ __forceinline__ SaaTetGeom saa_tet_geom(const int32_t *__restrict__ cell, const double *__restrict__ X)
{
    double P[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) P[a][i] = X[3 * (int64_t)cell[a] + i];
    double J[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) J[i][j] = P[j + 1][i] - P[0][i];
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    SaaTetGeom G;
    G.detJ = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double r = 1.0 / G.detJ;
    double inv[3][3];   // inv(J)
    inv[0][0] = c00 * r;
    inv[1][0] = c01 * r;
    inv[2][0] = c02 * r;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
    // dN/dxi = [[-1,-1,-1],[1,0,0],[0,1,0],[0,0,1]]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        G.g[0][i] = -(inv[0][i] + inv[1][i] + inv[2][i]);
        G.g[1][i] = inv[0][i];
        G.g[2][i] = inv[1][i];
        G.g[3][i] = inv[2][i];
    }
    return G;
}

// Sorted, de-duplicated neighbour list of one node from its incident elements.
//   inc_ptr / inc_slot: CSR node -> (4*element + corner) entries, ascending element order.
//   mode 0: counts[node] = number of neighbours;  mode 1: blk_col[blk_ptr[node] ...] = neighbours ascending.
__global__ void saa_k_asm_pattern(int64_t n_nodes, const int64_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc_slot,
                                  const int32_t *__restrict__ cells, int mode, int32_t *__restrict__ counts,
                                  const int64_t *__restrict__ blk_ptr, int32_t *__restrict__ blk_col, int *__restrict__ overflow)
{
    const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_nodes) return;
    int32_t nb[SAA_ASM_MAX_NB];
    int cnt = 0;
    for (int64_t k = inc_ptr[node]; k < inc_ptr[node + 1]; ++k) {
        const int32_t *cell = cells + 4 * (int64_t)(inc_slot[k] >> 2);
        for (int b = 0; b < 4; ++b) {
            const int32_t v = cell[b];
            int lo = 0, hi = cnt;                     // binary search in the sorted prefix
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (nb[mid] < v) lo = mid + 1; else hi = mid;
            }
            if (lo < cnt && nb[lo] == v) continue;
            if (cnt == SAA_ASM_MAX_NB) { *overflow = 1; continue; }
            for (int j = cnt; j > lo; --j) nb[j] = nb[j - 1];
            nb[lo] = v;
            ++cnt;
        }
    }
    if (mode == 0) {
        counts[node] = cnt;
    } else {
        int32_t *out = blk_col + blk_ptr[node];
        for (int j = 0; j < cnt; ++j) out[j] = nb[j];
    }
}

// Accumulate the 3x3 blocks of one node's rows: bval[(blk_ptr[node] + slot)*9 + 3*A + B].
__global__ void saa_k_asm_values(int64_t n_nodes, const int64_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc_slot,
                                 const int32_t *__restrict__ cells, const double *__restrict__ X, double lmd, double mu,
                                 const int64_t *__restrict__ blk_ptr, const int32_t *__restrict__ blk_col, double *__restrict__ bval)
{
    const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_nodes) return;
    const int64_t base = blk_ptr[node];
    const int nblk = (int)(blk_ptr[node + 1] - base);
    const double w = 0.25 / 6;                                                    // Qudrature.py:10
    for (int64_t k = inc_ptr[node]; k < inc_ptr[node + 1]; ++k) {                 // ascending element order
        const int32_t slot = inc_slot[k];
        const int a = slot & 3;
        const int32_t *cell = cells + 4 * (int64_t)(slot >> 2);
        const SaaTetGeom G = saa_tet_geom(cell, X);
        const double c = G.detJ * w;
        for (int b = 0; b < 4; ++b) {
            const int32_t v = cell[b];
            int lo = 0, hi = nblk;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (blk_col[base + mid] < v) lo = mid + 1; else hi = mid;
            }
            double *dst = bval + (base + lo) * 9;
            const double dotab = G.g[a][0] * G.g[b][0] + G.g[a][1] * G.g[b][1] + G.g[a][2] * G.g[b][2];
#pragma unroll
            for (int A = 0; A < 3; ++A)
#pragma unroll
                for (int B = 0; B < 3; ++B) {
                    double x = lmd * G.g[a][A] * G.g[b][B] + mu * G.g[a][B] * G.g[b][A];
                    if (A == B) x += mu * dotab;
                    const double kq = x * c;                                      // one quadrature addend (:112)
                    const double ke = ((kq + kq) + kq) + kq;                      // K starts at 0.0, four addends (:117)
                    dst[3 * A + B] = dst[3 * A + B] + ke;                         // K[P,Q] += Local_Ke[p,q] (:148)
                }
        }
    }
}

// Row lengths after dropping exact zeros (csr_matrix(dense), :150): row 3*node + A.
__global__ void saa_k_asm_row_nnz(int64_t n_nodes, const int64_t *__restrict__ blk_ptr, const double *__restrict__ bval,
                                  int64_t *__restrict__ row_nnz)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= 3 * n_nodes) return;
    const int64_t node = row / 3;
    const int A = (int)(row - 3 * node);
    int64_t cnt = 0;
    for (int64_t j = blk_ptr[node]; j < blk_ptr[node + 1]; ++j)
#pragma unroll
        for (int B = 0; B < 3; ++B) cnt += (bval[j * 9 + 3 * A + B] != 0.0) ? 1 : 0;
    row_nnz[row] = cnt;
}

__global__ void saa_k_asm_compact(int64_t n_nodes, const int64_t *__restrict__ blk_ptr, const int32_t *__restrict__ blk_col,
                                  const double *__restrict__ bval, const int64_t *__restrict__ indptr,
                                  int32_t *__restrict__ indices, double *__restrict__ data)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= 3 * n_nodes) return;
    const int64_t node = row / 3;
    const int A = (int)(row - 3 * node);
    int64_t o = indptr[row];
    for (int64_t j = blk_ptr[node]; j < blk_ptr[node + 1]; ++j)
#pragma unroll
        for (int B = 0; B < 3; ++B) {
            const double v = bval[j * 9 + 3 * A + B];
            if (v != 0.0) {
                indices[o] = 3 * blk_col[j] + B;
                data[o] = v;
                ++o;
            }
        }
}

// Partial lumped mass (one value per node) and un-ramped load vector of the LOCAL elements:
// Local_MKF (Mat_construction.py:36-73) row-summed (commons.py:103-107): m_a = sum_b sum_q N_a rho N_b detJ w,
// F_a[C] = sum_q N_a f_C detJ w, f = (0, -fz, -fz) (commons.py:35-38), element contributions in ascending order.
__global__ void saa_k_asm_mass_load(int64_t n_nodes, const int64_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc_slot,
                                    const int32_t *__restrict__ cells, const double *__restrict__ X, double rho, double fz,
                                    double *__restrict__ m_node, double *__restrict__ F)
{
    const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_nodes) return;
    const double q0 = 0.5854101966249685, q1 = 0.1381966011250105;               // Qudrature.py:8-9
    const double w = 0.25 / 6;
    double m = 0.0, fy = 0.0;
    for (int64_t k = inc_ptr[node]; k < inc_ptr[node + 1]; ++k) {
        const int32_t slot = inc_slot[k];
        const int a = slot & 3;
        const SaaTetGeom G = saa_tet_geom(cells + 4 * (int64_t)(slot >> 2), X);
        double me = 0.0, fe = 0.0;
        for (int q = 0; q < 4; ++q) {
            // shape functions at quadrature point q: N = (1-xi-eta-zeta, xi, eta, zeta); point q has q0 at coordinate q (q<3)
            double N[4];
            const double xi = (q == 0) ? q0 : q1, eta = (q == 1) ? q0 : q1, zeta = (q == 2) ? q0 : q1;
            N[0] = 1. - xi - eta - zeta; N[1] = xi; N[2] = eta; N[3] = zeta;
            for (int b = 0; b < 4; ++b) me = me + N[a] * rho * N[b] * G.detJ * w;
            fe = fe + N[a] * (-fz) * G.detJ * w;
        }
        m = m + me;
        fy = fy + fe;
    }
    m_node[node] = m;
    F[3 * node + 0] = 0.0;
    F[3 * node + 1] = fy;
    F[3 * node + 2] = fy;
}
