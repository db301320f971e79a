/*
 * oracle/fem_oracle.c — CPU restatement of the reference's explicit FE time step.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared with (tests/,
 * __graft_entry__.smoke(), and the cpu_baseline / --impl reference legs of bench.py).  It is never
 * linked, imported or executed by the product package.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement bit-for-bit against golden
 * vectors produced by the unmodified reference (oracle/gen_golden.py -> tests/golden/ fixtures) for
 * P = 1, 2, 3, 4, 8 partitions, up to 10 000 steps.
 *
 * Every operation is a separately rounded IEEE-754 binary64 operation in the reference's order;
 * compile with -ffp-contract=off (see oracle/Makefile) so that no FMA is formed.
 *
 * Reference lines restated (paths under /root/reference):
 *   Tools/Dynamic_solver.py:12     F_int = LocalK.dot(T.d0)  -> scipy.sparse _sparsetools.csr_matvec
 *                                  (third-party, unpinned; scipy 1.18.1 here): per row, sum starts at
 *                                  0.0 and adds data[j]*x[indices[j]] in stored (ascending column) order
 *   Tools/Dynamic_solver.py:13     F_ext = F_rankwise * linear_ramp(T.tn)   (commons.py:7-11)
 *   Tools/Dynamic_solver.py:17,29  central-difference update, Python left-to-right association
 *   Tools/Dynamic_solver.py:20,32  d1[Local_Dirichlet] = 0
 *   Tools/Distributed_tools.py:83-86,92  f_global = 0; for r in range(size): f_global[dofs_r] += f_r;
 *                                  return f_global[dofs_local]
 *   Data_prepare.py:223-235        loop, rotation d_n = d_0; d_0 = d1; tn = tn + dt
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n_dof;            /* 3 * local nodes */
    const int32_t *indptr;    /* LocalK (scipy CSR, int32) */
    const int32_t *indices;
    const double *data;
    const double *F;          /* F_rankwise (un-ramped) */
    const double *M;          /* l_M */
    int64_t n_dir;
    const int64_t *dir;       /* Local_Dirichlet (local DOF ids) */
    const int64_t *nodes;     /* Local_nodal_list (global node ids), n_dof/3 entries */
    double *d0, *dn, *d1, *fint;
} oracle_rank;

typedef struct {
    int size;
    int64_t n_global_nodes;   /* len(Points) */
    oracle_rank *ranks;
    double *f_global;         /* (3*len(Points)) scratch of syn_cpus */
    double tn;
} oracle_problem;

/* Dynamic_solver.py:12 -> scipy csr_matvec: separate multiply and add, stored order, start 0.0 */
void oracle_csr_matvec(int64_t n_row, const int32_t *indptr, const int32_t *indices, const double *data,
                       const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_row; ++i) {
        double sum = 0.0;
        for (int32_t jj = indptr[i]; jj < indptr[i + 1]; ++jj) {
            double prod = data[jj] * x[indices[jj]];
            sum = sum + prod;
        }
        y[i] = sum;
    }
}

/* commons.py:7-11 */
double oracle_linear_ramp(double t) { return (t <= 1) ? t : 1.0; }

/*
 * Dynamic_solver.py:17 / :29
 *   d1 = (T.dt**2*(F_ext - F_int) + 2*l_M*T.d0 - l_M*T.dn + T.dt/2*l_M*alpha*T.dn)/(l_M + 0.5*alpha*l_M*T.dt)
 * Python evaluates left to right:
 *   num = (((dt2*(Fe-Fi)) + ((2*M)*d0)) - (M*dn)) + ((((dt/2)*M)*alpha)*dn)
 *   den = M + (((0.5*alpha)*M)*dt)
 * dt2 = T.dt**2, dt_half = T.dt/2 and half_alpha = 0.5*alpha are scalars evaluated once by the
 * caller with the reference's own Python expressions and passed in.
 */
void oracle_cd_update(int64_t n, double dt2, double dt_half, double half_alpha, double alpha, double dt,
                      double ramp, const double *F, const double *Fi, const double *M, const double *d0,
                      const double *dn, double *d1)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double Fe = F[i] * ramp;                       /* :13 */
        double t1 = dt2 * (Fe - Fi[i]);
        double t2 = (2 * M[i]) * d0[i];
        double t3 = M[i] * dn[i];
        double t4 = ((dt_half * M[i]) * alpha) * dn[i];
        double num = ((t1 + t2) - t3) + t4;
        double den = M[i] + ((half_alpha * M[i]) * dt);
        d1[i] = num / den;
    }
}

/* Dynamic_solver.py:20,32 */
void oracle_dirichlet(double *d1, const int64_t *dir, int64_t n_dir)
{
    for (int64_t k = 0; k < n_dir; ++k) d1[dir[k]] = 0;
}

oracle_problem *oracle_problem_create(int size, int64_t n_global_nodes)
{
    oracle_problem *p = (oracle_problem *)calloc(1, sizeof(*p));
    p->size = size;
    p->n_global_nodes = n_global_nodes;
    p->ranks = (oracle_rank *)calloc((size_t)size, sizeof(oracle_rank));
    p->f_global = (double *)calloc((size_t)(3 * n_global_nodes), sizeof(double));
    p->tn = 0.0;                                        /* Data_prepare.py:215 */
    return p;
}

void oracle_problem_destroy(oracle_problem *p)
{
    if (!p) return;
    for (int r = 0; r < p->size; ++r) {
        free(p->ranks[r].d0); free(p->ranks[r].dn); free(p->ranks[r].d1); free(p->ranks[r].fint);
    }
    free(p->ranks); free(p->f_global); free(p);
}

/* arrays are borrowed (caller keeps them alive); state starts at d0 = dn = 0 (Data_prepare.py:171,178-189) */
void oracle_problem_set_rank(oracle_problem *p, int r, int64_t n_dof, const int32_t *indptr, const int32_t *indices,
                             const double *data, const double *F, const double *M, int64_t n_dir,
                             const int64_t *dir, const int64_t *nodes)
{
    oracle_rank *q = &p->ranks[r];
    q->n_dof = n_dof; q->indptr = indptr; q->indices = indices; q->data = data; q->F = F; q->M = M;
    q->n_dir = n_dir; q->dir = dir; q->nodes = nodes;
    q->d0 = (double *)calloc((size_t)n_dof, sizeof(double));
    q->dn = (double *)calloc((size_t)n_dof, sizeof(double));
    q->d1 = (double *)calloc((size_t)n_dof, sizeof(double));
    q->fint = (double *)calloc((size_t)n_dof, sizeof(double));
}

void oracle_problem_set_state(oracle_problem *p, int r, const double *d0, const double *dn, double tn)
{
    oracle_rank *q = &p->ranks[r];
    memcpy(q->d0, d0, sizeof(double) * (size_t)q->n_dof);
    memcpy(q->dn, dn, sizeof(double) * (size_t)q->n_dof);
    p->tn = tn;
}

void oracle_problem_get_state(const oracle_problem *p, int r, double *d0, double *dn, double *tn)
{
    const oracle_rank *q = &p->ranks[r];
    if (d0) memcpy(d0, q->d0, sizeof(double) * (size_t)q->n_dof);
    if (dn) memcpy(dn, q->dn, sizeof(double) * (size_t)q->n_dof);
    if (tn) *tn = p->tn;
}

/* Distributed_tools.py:83-86 and :92 — all ranks in-process */
static void oracle_syn_cpus(oracle_problem *p)
{
    memset(p->f_global, 0, sizeof(double) * (size_t)(3 * p->n_global_nodes));       /* :84 */
    for (int r = 0; r < p->size; ++r) {                                             /* :85 ascending rank */
        oracle_rank *q = &p->ranks[r];
        for (int64_t k = 0; k < q->n_dof / 3; ++k)
            for (int c = 0; c < 3; ++c)
                p->f_global[3 * q->nodes[k] + c] += q->fint[3 * k + c];             /* :86 */
    }
    for (int r = 0; r < p->size; ++r) {                                             /* :92 */
        oracle_rank *q = &p->ranks[r];
        for (int64_t k = 0; k < q->n_dof / 3; ++k)
            for (int c = 0; c < 3; ++c)
                q->fint[3 * k + c] = p->f_global[3 * q->nodes[k] + c];
    }
}

/*
 * n_steps iterations of Data_prepare.py:223-235 for all ranks.
 * model != 0 reproduces MODEL=True (Dynamic_solver.py:22): no synchronisation at all.
 * The scalars dt2 = dt**2, dt_half = dt/2, half_alpha = 0.5*alpha come from the caller (Python).
 */
void oracle_problem_run(oracle_problem *p, int64_t n_steps, double dt, double dt2, double dt_half,
                        double half_alpha, double alpha, int model)
{
    for (int64_t s = 0; s < n_steps; ++s) {
        double ramp = oracle_linear_ramp(p->tn);
        for (int r = 0; r < p->size; ++r) {
            oracle_rank *q = &p->ranks[r];
            oracle_csr_matvec(q->n_dof, q->indptr, q->indices, q->data, q->d0, q->fint);   /* :12 */
        }
        if (p->size != 1 && !model) oracle_syn_cpus(p);                                    /* :25-26 */
        for (int r = 0; r < p->size; ++r) {
            oracle_rank *q = &p->ranks[r];
            oracle_cd_update(q->n_dof, dt2, dt_half, half_alpha, alpha, dt, ramp, q->F, q->fint, q->M,
                             q->d0, q->dn, q->d1);                                         /* :17 / :29 */
            oracle_dirichlet(q->d1, q->dir, q->n_dir);                                     /* :20 / :32 */
            double *old = q->dn;                                                           /* :233-234 */
            q->dn = q->d0; q->d0 = q->d1; q->d1 = old;
        }
        p->tn = p->tn + dt;                                                                /* :235 */
    }
}
