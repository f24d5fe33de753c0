/*
 * pyro_oracle.c -- CPU restatement of the two pyro2 hot paths.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path in pyro2_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product never links, imports or calls it.
 *
 * It restates, stage by stage and full-array like the reference (one pass per operator, with
 * the reference's operation order, no FMA contraction: build with -ffp-contract=off):
 *   HP-1  pyro/compressible/simulation.py:290-450 (evolve), :49-102 (cons<->prim), :267-288 (dt)
 *         pyro/compressible/unsplit_fluxes.py:134-244, 333-494, 497-549
 *         pyro/compressible/interface.py:6-236 (states), :240-378 (artificial_viscosity)
 *         pyro/compressible/riemann.py:597-678 (estimate_wave_speed), :682-860 (riemann_hllc),
 *                                      :1105-1179 (consFlux)
 *         pyro/mesh/reconstruction.py:9-183 (limit, limit2, limit4, flatten, flatten_multid)
 *         pyro/mesh/array_indexer.py:150-274 (fill_ghost), :98-111 (norm)
 *   HP-2  pyro/multigrid/MG.py:529-542 (_compute_residual), :544-599 (smooth), :623-697 (solve),
 *         :699-778 (v_cycle);  pyro/mesh/patch.py:640-676 (restrict), :678-736 (prolong)
 *
 * Parity pin: validated against the unmodified reference imported in the build container
 * (tests/test_oracle_vs_reference.py) and against fixtures it generated (tests/golden/).
 *
 * Layout: SoA planes a[n][i][j], j (y) contiguous; x = axis i (slowest).  The reference stores
 * [i][j][n]; the test harness transposes.  Index conventions as SURVEY.md 9.1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j) ((size_t)(i) * (size_t)qy + (size_t)(j))

enum { BC_OUTFLOW = 0, BC_REFLECT_EVEN = 1, BC_REFLECT_ODD = 2, BC_PERIODIC = 3 };
/* conserved order (compressible/simulation.py:223-226): density, energy, x-momentum, y-momentum */
enum { IDENS = 0, IENER = 1, IXMOM = 2, IYMOM = 3 };
enum { IRHO = 0, IU = 1, IV = 2, IP = 3 };

/* ------------------------------------------------------------------------------------------
 * ghost fill -- array_indexer.py:150-274.  x faces first, then y faces over the full x range
 * (corners inherit).  xl_val.. are the optional inhomogeneous boundary values (length qy / qx)
 * which only touch the first ghost cell (array_indexer.py:166-183).
 * ---------------------------------------------------------------------------------------- */
#define DEFINE_FILL_GHOST(NAME, T)                                                              \
    void NAME(T *a, int nx, int ny, int ng, int xlb, int xrb, int ylb, int yrb,                 \
              const T *xl_val, const T *xr_val, const T *yl_val, const T *yr_val, double dx,    \
              double dy)                                                                        \
    {                                                                                           \
        const int qx = nx + 2 * ng, qy = ny + 2 * ng;                                           \
        const int ilo = ng, ihi = ng + nx - 1, jlo = ng, jhi = ng + ny - 1;                     \
        int i, j;                                                                               \
        /* -x */                                                                                \
        if (xlb == BC_OUTFLOW) {                                                                \
            if (!xl_val) { for (i = 0; i < ilo; i++) for (j = 0; j < qy; j++) a[IDX(i, j)] = a[IDX(ilo, j)]; } \
            else for (j = 0; j < qy; j++) a[IDX(ilo - 1, j)] = a[IDX(ilo, j)] - (T)(dx * xl_val[j]); \
        } else if (xlb == BC_REFLECT_EVEN) {                                                    \
            for (i = 0; i < ilo; i++) for (j = 0; j < qy; j++) a[IDX(i, j)] = a[IDX(2 * ng - i - 1, j)]; \
        } else if (xlb == BC_REFLECT_ODD) {                                                     \
            if (!xl_val) { for (i = 0; i < ilo; i++) for (j = 0; j < qy; j++) a[IDX(i, j)] = -a[IDX(2 * ng - i - 1, j)]; } \
            else for (j = 0; j < qy; j++) a[IDX(ilo - 1, j)] = 2 * xl_val[j] - a[IDX(ilo, j)];  \
        } else if (xlb == BC_PERIODIC) {                                                        \
            for (i = 0; i < ilo; i++) for (j = 0; j < qy; j++) a[IDX(i, j)] = a[IDX(ihi - ng + i + 1, j)]; \
        }                                                                                       \
        /* +x */                                                                                \
        if (xrb == BC_OUTFLOW) {                                                                \
            if (!xr_val) { for (i = ihi + 1; i < qx; i++) for (j = 0; j < qy; j++) a[IDX(i, j)] = a[IDX(ihi, j)]; } \
            else for (j = 0; j < qy; j++) a[IDX(ihi + 1, j)] = a[IDX(ihi, j)] + (T)(dx * xr_val[j]); \
        } else if (xrb == BC_REFLECT_EVEN) {                                                    \
            for (i = 0; i < ng; i++) for (j = 0; j < qy; j++) a[IDX(ihi + 1 + i, j)] = a[IDX(ihi - i, j)]; \
        } else if (xrb == BC_REFLECT_ODD) {                                                     \
            if (!xr_val) { for (i = 0; i < ng; i++) for (j = 0; j < qy; j++) a[IDX(ihi + 1 + i, j)] = -a[IDX(ihi - i, j)]; } \
            else for (j = 0; j < qy; j++) a[IDX(ihi + 1, j)] = 2 * xr_val[j] - a[IDX(ihi, j)];  \
        } else if (xrb == BC_PERIODIC) {                                                        \
            for (i = ihi + 1; i < qx; i++) for (j = 0; j < qy; j++) a[IDX(i, j)] = a[IDX(i - ihi - 1 + ng, j)]; \
        }                                                                                       \
        /* -y */                                                                                \
        if (ylb == BC_OUTFLOW) {                                                                \
            if (!yl_val) { for (i = 0; i < qx; i++) for (j = 0; j < jlo; j++) a[IDX(i, j)] = a[IDX(i, jlo)]; } \
            else for (i = 0; i < qx; i++) a[IDX(i, jlo - 1)] = a[IDX(i, jlo)] - (T)(dy * yl_val[i]); \
        } else if (ylb == BC_REFLECT_EVEN) {                                                    \
            for (i = 0; i < qx; i++) for (j = 0; j < jlo; j++) a[IDX(i, j)] = a[IDX(i, 2 * ng - j - 1)]; \
        } else if (ylb == BC_REFLECT_ODD) {                                                     \
            if (!yl_val) { for (i = 0; i < qx; i++) for (j = 0; j < jlo; j++) a[IDX(i, j)] = -a[IDX(i, 2 * ng - j - 1)]; } \
            else for (i = 0; i < qx; i++) a[IDX(i, jlo - 1)] = 2 * yl_val[i] - a[IDX(i, jlo)];  \
        } else if (ylb == BC_PERIODIC) {                                                        \
            for (i = 0; i < qx; i++) for (j = 0; j < jlo; j++) a[IDX(i, j)] = a[IDX(i, jhi - ng + j + 1)]; \
        }                                                                                       \
        /* +y */                                                                                \
        if (yrb == BC_OUTFLOW) {                                                                \
            if (!yr_val) { for (i = 0; i < qx; i++) for (j = jhi + 1; j < qy; j++) a[IDX(i, j)] = a[IDX(i, jhi)]; } \
            else for (i = 0; i < qx; i++) a[IDX(i, jhi + 1)] = a[IDX(i, jhi)] + (T)(dy * yr_val[i]); \
        } else if (yrb == BC_REFLECT_EVEN) {                                                    \
            for (i = 0; i < qx; i++) for (j = 0; j < ng; j++) a[IDX(i, jhi + 1 + j)] = a[IDX(i, jhi - j)]; \
        } else if (yrb == BC_REFLECT_ODD) {                                                     \
            if (!yr_val) { for (i = 0; i < qx; i++) for (j = 0; j < ng; j++) a[IDX(i, jhi + 1 + j)] = -a[IDX(i, jhi - j)]; } \
            else for (i = 0; i < qx; i++) a[IDX(i, jhi + 1)] = 2 * yr_val[i] - a[IDX(i, jhi)];  \
        } else if (yrb == BC_PERIODIC) {                                                        \
            for (i = 0; i < qx; i++) for (j = jhi + 1; j < qy; j++) a[IDX(i, j)] = a[IDX(i, j - jhi - 1 + ng)]; \
        }                                                                                       \
    }

DEFINE_FILL_GHOST(orc_fill_ghost_f64, double)
DEFINE_FILL_GHOST(orc_fill_ghost_i64, int64_t)

/* ------------------------------------------------------------------------------------------
 * CFL timestep -- compressible/simulation.py:267-288 + derives.py:6-69.  Minimum over the FULL
 * array including ghosts.  Returns cfl*min(...).
 * ---------------------------------------------------------------------------------------- */
double orc_cfl_dt(const double *U, int nx, int ny, int ng, double dx, double dy, double gamma,
                  double cfl)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    const size_t np = (size_t)qx * qy;
    const double *dens = U + IDENS * np, *ener = U + IENER * np, *xmom = U + IXMOM * np,
                 *ymom = U + IYMOM * np;
    double xmin = INFINITY, ymin = INFINITY;
#pragma omp parallel for reduction(min : xmin, ymin)
    for (size_t k = 0; k < np; k++) {
        double u = xmom[k] / dens[k];
        double v = ymom[k] / dens[k];
        double e = (ener[k] - 0.5 * dens[k] * (u * u + v * v)) / dens[k];
        double p = dens[k] * e * (gamma - 1.0);
        double cs = sqrt(gamma * p / dens[k]);
        double xt = dx / (fabs(u) + cs);
        double yt = dy / (fabs(v) + cs);
        if (xt < xmin) xmin = xt;
        if (yt < ymin) ymin = yt;
    }
    return cfl * (xmin < ymin ? xmin : ymin);
}

/* ------------------------------------------------------------------------------------------
 * compressible step
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double gamma;
    double z0, z1, delta; /* flattening (compressible/_defaults) */
    double cvisc;
    int limiter;        /* 0 none, 1 MC 2nd, 2 MC 4th */
    int use_flattening; /* 0/1 */
    int no_avisc_xhi;   /* 1 (default): reproduce SURVEY 9.2-13, avisco_x not set on the +x face */
    int no_avisc_yhi;
    /* gravity along -y (compressible.grav); src_bc = BC codes (xl, xr, yl, yr) of the four source arrays
       dens_src, E_src, xmom_src, ymom_src in VARIABLE-INDEX order (dens, ener, xmom, ymom); the "hse" type
       copies the first interior row like outflow (compressible/BC.py:55-63, 111-117) */
    double grav;
    int src_bc[16];
    /* compressible.riemann: 0 HLLC, 1 CGF, 2 HLLC_lm; xl_solid / yl_solid: the -x / -y boundary is a solid wall
       (boundary.bc_is_solid), which CGF uses to zero the normal velocity at that face */
    int riemann, xl_solid, yl_solid;
    /* problem heating source S_ener = dens * heat_rate * heat_profile[i, j] (compressible/problems/heating.py,
       plume.py, convection.py: source_terms); heat_profile = NULL: none.  Its ghost cells are irrelevant: the
       reference evaluates the source on the ghost-filled state and then ghost-fills the source arrays */
    double heat_rate;
    const double *heat_profile;
    /* sponge (simulation.py:164-184, 425-441): do_sponge, rho_begin, rho_full, timescale */
    int do_sponge;
    double sponge_rho_begin, sponge_rho_full, sponge_timescale;
    /* SphericalPolar geometry (mesh/patch.py:242-312; x = r, y = theta), NULL = Cartesian: full (qx, qy) planes
       computed by the caller with the reference's numpy expressions (oracle.py: spherical_geometry) */
    const struct orc_geom *geom;
} orc_comp_params;

typedef struct orc_geom {
    double xmin, ymin;
    const double *Ly, *Ax, *Ay, *V, *dlogAx, *dlogAy, *x2d;      /* Lx = dx everywhere */
} orc_geom;

/* optional per-stage dumps, each (4 or 1) planes of qx*qy doubles; NULL = skip */
typedef struct {
    double *q, *xi, *ldx, *ldy;
    double *Uxl_hat, *Uxr_hat, *Uyl_hat, *Uyr_hat; /* after prim_to_cons, before transverse */
    double *Fx_t, *Fy_t;                           /* transverse fluxes */
    double *Uxl, *Uxr, *Uyl, *Uyr;                 /* after transverse correction */
    double *Fx, *Fy;                               /* final fluxes incl. artificial viscosity */
} orc_comp_stages;


/* max / min with the semantics of Python's built-ins, which numba compiles the reference's scalar max() / min()
 * calls to: the SECOND argument wins only if it compares greater (less), so max(a, nan) = a and max(nan, b) = nan
 * (C's fmax / fmin would return the non-NaN operand in both cases).  It matters where the reference feeds an
 * unphysical interface state to a Riemann solver -- c = max(smallc, sqrt(negative)) = smallc -- and carries on. */
static inline double pymax(double a, double b) { return b > a ? b : a; }
static inline double pymin(double a, double b) { return b < a ? b : a; }
static double *zalloc(size_t n) { return (double *)calloc(n, sizeof(double)); }

/* simulation.py:49-80 */
static int cons_to_prim(const double *U, double *q, int qx, int qy, int ng, double gamma)
{
    const size_t np = (size_t)qx * qy;
    int bad = 0;
#pragma omp parallel for reduction(| : bad)
    for (int i = 0; i < qx; i++)
        for (int j = 0; j < qy; j++) {
            size_t k = IDX(i, j);
            double rho = U[IDENS * np + k];
            double u = 0.0, v = 0.0, e = 0.0;
            if (rho != 0.0) {
                u = U[IXMOM * np + k] / rho;
                v = U[IYMOM * np + k] / rho;
            }
            if (rho != 0.0) e = (U[IENER * np + k] - 0.5 * rho * (u * u + v * v)) / rho;
            q[IRHO * np + k] = rho;
            q[IU * np + k] = u;
            q[IV * np + k] = v;
            q[IP * np + k] = rho * e * (gamma - 1.0);
            if (i >= ng && i < qx - ng && j >= ng && j < qy - ng && !(e > 0.0 && rho > 0.0)) bad = 1;
        }
    return bad;
}

/* simulation.py:83-102 */
static void prim_to_cons(const double *q, double *U, int qx, int qy, double gamma)
{
    const size_t np = (size_t)qx * qy;
#pragma omp parallel for
    for (size_t k = 0; k < np; k++) {
        double rho = q[IRHO * np + k], u = q[IU * np + k], v = q[IV * np + k], p = q[IP * np + k];
        U[IDENS * np + k] = rho;
        U[IXMOM * np + k] = u * rho;
        U[IYMOM * np + k] = v * rho;
        double rhoe = p / (gamma - 1.0);
        U[IENER * np + k] = rhoe + 0.5 * rho * (u * u + v * v);
    }
}

/* reconstruction.py:123-164; idir 1 = x */
static void flatten1d(const double *q, double *xi, int qx, int qy, int ng, int idir, double z0,
                      double z1, double delta)
{
    const size_t np = (size_t)qx * qy;
    const double *p = q + IP * np;
    const double *un = q + (idir == 1 ? IU : IV) * np;
    const double smallp = 1.e-10;
    const int di = idir == 1 ? 1 : 0, dj = idir == 1 ? 0 : 1;
#pragma omp parallel for
    for (int i = 0; i < qx; i++)
        for (int j = 0; j < qy; j++) {
            int in = (i >= ng - 2 && i <= qx - ng + 1 && j >= ng - 2 && j <= qy - ng + 1);
            double t1 = 0.0, t2 = 0.0, t2b = 0.0, t1b = 0.0;
            if (in) {
                t1 = fabs(p[IDX(i + di, j + dj)] - p[IDX(i - di, j - dj)]);
                t2 = fabs(p[IDX(i + 2 * di, j + 2 * dj)] - p[IDX(i - 2 * di, j - 2 * dj)]);
            }
            double z = t1 / pymax(t2, smallp);
            if (in) {
                t2b = t1 / pymin(p[IDX(i + di, j + dj)], p[IDX(i - di, j - dj)]);
                t1b = un[IDX(i - di, j - dj)] - un[IDX(i + di, j + dj)];
            }
            double x = pymin(1.0, pymax(0.0, 1.0 - (z - z0) / (z1 - z0)));
            xi[IDX(i, j)] = (t1b > 0.0 && t2b > delta) ? x : 1.0;
        }
}

/* reconstruction.py:167-183 */
static void flatten_multid(const double *q, const double *xi_x, const double *xi_y, double *xi,
                           int qx, int qy, int ng)
{
    const size_t np = (size_t)qx * qy;
    const double *p = q + IP * np;
    memset(xi, 0, np * sizeof(double));
#pragma omp parallel for
    for (int i = ng - 2; i <= qx - ng + 1; i++)
        for (int j = ng - 2; j <= qy - ng + 1; j++) {
            double px = (p[IDX(i + 1, j)] - p[IDX(i - 1, j)] > 0) ? xi_x[IDX(i - 1, j)] : xi_x[IDX(i + 1, j)];
            double py = (p[IDX(i, j + 1)] - p[IDX(i, j - 1)] > 0) ? xi_y[IDX(i, j - 1)] : xi_y[IDX(i, j + 1)];
            xi[IDX(i, j)] = pymin(pymin(xi_x[IDX(i, j)], px), pymin(xi_y[IDX(i, j)], py));
        }
}

/* shared tail of limit2/limit4: reconstruction.py:87-89 / 116-118 */
static inline double mc_select(double dc, double dl, double dr)
{
    double d1 = 2.0 * (fabs(dl) < fabs(dr) ? dl : dr);
    double dt = fabs(dc) < fabs(d1) ? dc : d1;
    return (dl * dr > 0.0) ? dt : 0.0;
}

/* reconstruction.py:69-91 (limit2), :58-66 (nolimit); a is one plane, result zero outside buf=2 */
static void limit2(const double *a, double *lda, int qx, int qy, int ng, int idir, int nolimit)
{
    const int di = idir == 1 ? 1 : 0, dj = idir == 1 ? 0 : 1;
    memset(lda, 0, (size_t)qx * qy * sizeof(double));
#pragma omp parallel for
    for (int i = ng - 2; i <= qx - ng + 1; i++)
        for (int j = ng - 2; j <= qy - ng + 1; j++) {
            double ap = a[IDX(i + di, j + dj)], am = a[IDX(i - di, j - dj)], a0 = a[IDX(i, j)];
            double dc = 0.5 * (ap - am);
            lda[IDX(i, j)] = nolimit ? dc : mc_select(dc, ap - a0, a0 - am);
        }
}

/* reconstruction.py:94-120 */
static void limit4(const double *a, double *lda, double *tmp, int qx, int qy, int ng, int idir)
{
    const int di = idir == 1 ? 1 : 0, dj = idir == 1 ? 0 : 1;
    limit2(a, tmp, qx, qy, ng, idir, 0);
    memset(lda, 0, (size_t)qx * qy * sizeof(double));
#pragma omp parallel for
    for (int i = ng - 2; i <= qx - ng + 1; i++)
        for (int j = ng - 2; j <= qy - ng + 1; j++) {
            double ap = a[IDX(i + di, j + dj)], am = a[IDX(i - di, j - dj)], a0 = a[IDX(i, j)];
            double dc = (2. / 3.) * (ap - am - 0.25 * (tmp[IDX(i + di, j + dj)] + tmp[IDX(i - di, j - dj)]));
            lda[IDX(i, j)] = mc_select(dc, ap - a0, a0 - am);
        }
}

/* interface.py:6-236 (Cartesian: dloga = 0 so the geometric source vanishes) */
static void trace_states(int idir, const double *qv, const double *dqv, double *q_l, double *q_r,
                         int qx, int qy, int ng, double dx, double dt, double gamma, const double *Larr,
                         const double *dloga)
{
    const size_t np = (size_t)qx * qy;
    const int nx = qx - 2 * ng, ny = qy - 2 * ng;
    const int ilo = ng, ihi = ng + nx, jlo = ng, jhi = ng + ny;
    const double dtdx_c = dt / dx;
    const int in = idir == 1 ? IU : IV; /* normal velocity slot */
    const int it = idir == 1 ? IV : IU; /* transverse velocity slot */
    memset(q_l, 0, 4 * np * sizeof(double));
    memset(q_r, 0, 4 * np * sizeof(double));
#pragma omp parallel for
    for (int i = ilo - 2; i < ihi + 2; i++)
        for (int j = jlo - 2; j < jhi + 2; j++) {
            size_t k = IDX(i, j);
            size_t kl = idir == 1 ? IDX(i + 1, j) : IDX(i, j + 1);
            /* dtdx = dt / dx[i, j]: the cell's own length along idir (SphericalPolar: Ly = r dtheta varies with i) */
            const double dtdx = Larr ? dt / Larr[k] : dtdx_c, dtdx4 = 0.25 * dtdx;
            double q[4], dq[4], lvec[4][4], rvec[4][4], e_val[4], betal[4], betar[4];
            for (int m = 0; m < 4; m++) { q[m] = qv[m * np + k]; dq[m] = dqv[m * np + k]; }
            double cs = sqrt(gamma * q[IP] / q[IRHO]);
            memset(lvec, 0, sizeof lvec);
            memset(rvec, 0, sizeof rvec);
            e_val[0] = q[in] - cs; e_val[1] = q[in]; e_val[2] = q[in]; e_val[3] = q[in] + cs;
            lvec[0][in] = -0.5 * q[IRHO] / cs; lvec[0][IP] = 0.5 / (cs * cs);
            lvec[1][IRHO] = 1.0;               lvec[1][IP] = -1.0 / (cs * cs);
            lvec[2][it] = 1.0;
            lvec[3][in] = 0.5 * q[IRHO] / cs;  lvec[3][IP] = 0.5 / (cs * cs);
            rvec[0][IRHO] = 1.0; rvec[0][in] = -cs / q[IRHO]; rvec[0][IP] = cs * cs;
            rvec[1][IRHO] = 1.0;
            rvec[2][it] = 1.0;
            rvec[3][IRHO] = 1.0; rvec[3][in] = cs / q[IRHO];  rvec[3][IP] = cs * cs;

            double factor = 0.5 * (1.0 - dtdx * pymax(e_val[3], 0.0));
            double ql[4], qr[4];
            for (int m = 0; m < 4; m++) ql[m] = q[m] + factor * dq[m];
            factor = 0.5 * (1.0 + dtdx * pymin(e_val[0], 0.0));
            for (int m = 0; m < 4; m++) qr[m] = q[m] - factor * dq[m];

            for (int m = 0; m < 4; m++) {
                double asum = 0.0;
                for (int n = 0; n < 4; n++) asum += lvec[m][n] * dq[n];
                betal[m] = dtdx4 * (e_val[3] - e_val[m]) * (copysign(1.0, e_val[m]) + 1.0) * asum;
                betar[m] = dtdx4 * (e_val[0] - e_val[m]) * (1.0 - copysign(1.0, e_val[m])) * asum;
            }
            for (int m = 0; m < 4; m++) {
                double sum_l = 0.0, sum_r = 0.0;
                for (int n = 0; n < 4; n++) { sum_l += betal[n] * rvec[n][m]; sum_r += betar[n] * rvec[n][m]; }
                q_l[m * np + kl] = ql[m] + sum_l;
                q_r[m * np + k] = qr[m] + sum_r;
            }
            if (dloga) {
                /* geometric source of the divergence in curvilinear coordinates (interface.py:218-234) */
                const double rho_source = -0.5 * dt * dloga[k] * q[IRHO] * q[in];
                q_l[IRHO * np + kl] += rho_source;
                q_r[IRHO * np + k] += rho_source;
                q_l[IP * np + kl] += rho_source * cs * cs;
                q_r[IP * np + k] += rho_source * cs * cs;
            }
        }
}

/* riemann.py:597-678 */
static void estimate_wave_speed(double rho_l, double u_l, double p_l, double c_l, double rho_r,
                                double u_r, double p_r, double c_r, double gamma, double *S_l,
                                double *S_r)
{
    double p_max = pymax(p_l, p_r), p_min = pymin(p_l, p_r);
    double Q = p_max / p_min;
    double rho_avg = 0.5 * (rho_l + rho_r), c_avg = 0.5 * (c_l + c_r);
    double factor = rho_avg * c_avg;
    double pstar = 0.5 * (p_l + p_r) + 0.5 * (u_l - u_r) * factor;
    double ustar = 0.5 * (u_l + u_r) + 0.5 * (p_l - p_r) / factor;
    if (Q > 2 && (pstar < p_min || pstar > p_max)) {
        if (pstar < p_min) {
            double z = (gamma - 1.0) / (2.0 * gamma);
            double p_lr = pow(p_l / p_r, z);
            ustar = (p_lr * u_l / c_l + u_r / c_r + 2.0 * (p_lr - 1.0) / (gamma - 1.0)) /
                    (p_lr / c_l + 1.0 / c_r);
            pstar = 0.5 * (p_l * pow(1.0 + (gamma - 1.0) * (u_l - ustar) / (2.0 * c_l), 1.0 / z) +
                           p_r * pow(1.0 + (gamma - 1.0) * (ustar - u_r) / (2.0 * c_r), 1.0 / z));
        } else {
            double A_r = 2.0 / ((gamma + 1.0) * rho_r), B_r = p_r * (gamma - 1.0) / (gamma + 1.0);
            double A_l = 2.0 / ((gamma + 1.0) * rho_l), B_l = p_l * (gamma - 1.0) / (gamma + 1.0);
            double p_guess = pymax(0.0, pstar);
            double g_l = sqrt(A_l / (p_guess + B_l)), g_r = sqrt(A_r / (p_guess + B_r));
            pstar = (g_l * p_l + g_r * p_r - (u_r - u_l)) / (g_l + g_r);
            ustar = 0.5 * (u_l + u_r) + 0.5 * ((pstar - p_r) * g_r - (pstar - p_l) * g_l);
        }
    }
    (void)ustar;
    if (pstar <= p_l) *S_l = u_l - c_l;
    else *S_l = u_l - c_l * sqrt(1.0 + ((gamma + 1.0) / (2.0 * gamma)) * (pstar / p_l - 1.0));
    if (pstar <= p_r) *S_r = u_r + c_r;
    else /* (gamma+1)/(2/gamma): reference quirk, SURVEY 9.2-1 (riemann.py:675) */
        *S_r = u_r + c_r * sqrt(1.0 + ((gamma + 1.0) / (2.0 / gamma)) * (pstar / p_r - 1.0));
}

/* riemann.py:1105-1179, 1-d state branch, Cartesian */
static void cons_flux_geom(int idir, double gamma, const double U[4], double F[4], int coord_type);
static void cons_flux(int idir, double gamma, const double U[4], double F[4]) { cons_flux_geom(idir, gamma, U, F, 0); }
/* coord_type 1 (SphericalPolar): the pressure is left out of the momentum flux, its gradient is applied
   separately (riemann.py:1156-1158, 1171-1173) */
static void cons_flux_geom(int idir, double gamma, const double U[4], double F[4], int coord_type)
{
    double u = 0.0, v = 0.0;
    if (U[IDENS] != 0.0) { u = U[IXMOM] / U[IDENS]; v = U[IYMOM] / U[IDENS]; }
    double p = (U[IENER] - 0.5 * U[IDENS] * (u * u + v * v)) * (gamma - 1.0);
    if (idir == 1) {
        F[IDENS] = U[IDENS] * u;
        F[IXMOM] = U[IXMOM] * u;
        if (coord_type == 0) F[IXMOM] += p;
        F[IYMOM] = U[IYMOM] * u;
        F[IENER] = (U[IENER] + p) * u;
    } else {
        F[IDENS] = U[IDENS] * v;
        F[IXMOM] = U[IXMOM] * v;
        F[IYMOM] = U[IYMOM] * v;
        if (coord_type == 0) F[IYMOM] += p;
        F[IENER] = (U[IENER] + p) * v;
    }
}

/* riemann.py:682-860 */
static void riemann_hllc(int idir, const double *U_l, const double *U_r, double *F, int qx, int qy,
                         int ng, double gamma)
{
    const size_t np = (size_t)qx * qy;
    const int nx = qx - 2 * ng, ny = qy - 2 * ng;
    const int ilo = ng, ihi = ng + nx, jlo = ng, jhi = ng + ny;
    const double smallc = 1.e-10, smallp = 1.e-10;
    const int imn = idir == 1 ? IXMOM : IYMOM, imt = idir == 1 ? IYMOM : IXMOM;
    memset(F, 0, 4 * np * sizeof(double));
#pragma omp parallel for
    for (int i = ilo - 1; i < ihi + 1; i++)
        for (int j = jlo - 1; j < jhi + 1; j++) {
            size_t k = IDX(i, j);
            double Ul[4], Ur[4], Us[4], Fk[4];
            for (int m = 0; m < 4; m++) { Ul[m] = U_l[m * np + k]; Ur[m] = U_r[m * np + k]; }
            double rho_l = Ul[IDENS];
            double un_l = Ul[imn] / rho_l, ut_l = Ul[imt] / rho_l;
            double rhoe_l = Ul[IENER] - 0.5 * rho_l * (un_l * un_l + ut_l * ut_l);
            double p_l = pymax(rhoe_l * (gamma - 1.0), smallp);
            double rho_r = Ur[IDENS];
            double un_r = Ur[imn] / rho_r, ut_r = Ur[imt] / rho_r;
            double rhoe_r = Ur[IENER] - 0.5 * rho_r * (un_r * un_r + ut_r * ut_r);
            double p_r = pymax(rhoe_r * (gamma - 1.0), smallp);
            double c_l = pymax(smallc, sqrt(gamma * p_l / rho_l));
            double c_r = pymax(smallc, sqrt(gamma * p_r / rho_r));
            double S_l, S_r;
            estimate_wave_speed(rho_l, un_l, p_l, c_l, rho_r, un_r, p_r, c_r, gamma, &S_l, &S_r);
            double S_c = (p_r - p_l + rho_l * un_l * (S_l - un_l) - rho_r * un_r * (S_r - un_r)) /
                         (rho_l * (S_l - un_l) - rho_r * (S_r - un_r));
            if (S_r <= 0.0) {
                cons_flux(idir, gamma, Ur, Fk);
            } else if (S_c <= 0.0 && 0.0 < S_r) {
                double f = rho_r * (S_r - un_r) / (S_r - S_c);
                Us[IDENS] = f;
                Us[imn] = f * S_c;
                Us[imt] = f * ut_r;
                Us[IENER] = f * (Ur[IENER] / rho_r + (S_c - un_r) * (S_c + p_r / (rho_r * (S_r - un_r))));
                cons_flux(idir, gamma, Ur, Fk);
                for (int m = 0; m < 4; m++) Fk[m] = Fk[m] + S_r * (Us[m] - Ur[m]);
            } else if (S_l < 0.0 && 0.0 < S_c) {
                double f = rho_l * (S_l - un_l) / (S_l - S_c);
                Us[IDENS] = f;
                Us[imn] = f * S_c;
                Us[imt] = f * ut_l;
                Us[IENER] = f * (Ul[IENER] / rho_l + (S_c - un_l) * (S_c + p_l / (rho_l * (S_l - un_l))));
                cons_flux(idir, gamma, Ul, Fk);
                for (int m = 0; m < 4; m++) Fk[m] = Fk[m] + S_l * (Us[m] - Ul[m]);
            } else {
                cons_flux(idir, gamma, Ul, Fk);
            }
            for (int m = 0; m < 4; m++) F[m * np + k] = Fk[m];
        }
}

/* riemann.py:9-310 (riemann_cgf: the two-shock solver of Colella, Glaz & Ferguson) followed by consFlux
 * (riemann_flux :1075-1085); lower_solid: the normal velocity at the face on a solid lower boundary is zero
 * (the upper test `i == ihi + 1` in the reference can never fire inside its loop range) */
static void riemann_cgf_geom(int idir, const double *U_l, const double *U_r, double *F, int qx, int qy, int ng,
                             double gamma, int lower_solid, int coord_type, double *Ustate);
static void riemann_cgf(int idir, const double *U_l, const double *U_r, double *F, int qx, int qy, int ng,
                        double gamma, int lower_solid)
{
    riemann_cgf_geom(idir, U_l, U_r, F, qx, qy, ng, gamma, lower_solid, 0, NULL);
}
/* Ustate (optional, 4 planes): the conserved interface state itself (riemann_flux(..., return_cons=True),
   riemann.py:1097-1099), whose pressure the SphericalPolar update differences */
static void riemann_cgf_geom(int idir, const double *U_l, const double *U_r, double *F, int qx, int qy, int ng,
                             double gamma, int lower_solid, int coord_type, double *Ustate)
{
    const size_t np = (size_t)qx * qy;
    const int nx = qx - 2 * ng, ny = qy - 2 * ng;
    const int ilo = ng, ihi = ng + nx, jlo = ng, jhi = ng + ny;
    const double smallc = 1.e-10, smallrho = 1.e-10, smallp = 1.e-10;
    const int imn = idir == 1 ? IXMOM : IYMOM, imt = idir == 1 ? IYMOM : IXMOM;
    memset(F, 0, 4 * np * sizeof(double));
#pragma omp parallel for
    for (int i = ilo - 1; i < ihi + 1; i++)
        for (int j = jlo - 1; j < jhi + 1; j++) {
            const size_t k = IDX(i, j);
            const double rho_l = U_l[IDENS * np + k];
            const double un_l = U_l[imn * np + k] / rho_l, ut_l = U_l[imt * np + k] / rho_l;
            const double rhoe_l = U_l[IENER * np + k] - 0.5 * rho_l * (un_l * un_l + ut_l * ut_l);
            const double p_l = pymax(rhoe_l * (gamma - 1.0), smallp);
            const double rho_r = U_r[IDENS * np + k];
            const double un_r = U_r[imn * np + k] / rho_r, ut_r = U_r[imt * np + k] / rho_r;
            const double rhoe_r = U_r[IENER * np + k] - 0.5 * rho_r * (un_r * un_r + ut_r * ut_r);
            const double p_r = pymax(rhoe_r * (gamma - 1.0), smallp);
            const double W_l = pymax(smallrho * smallc, sqrt(gamma * p_l * rho_l));
            const double W_r = pymax(smallrho * smallc, sqrt(gamma * p_r * rho_r));
            const double c_l = pymax(smallc, sqrt(gamma * p_l / rho_l));
            const double c_r = pymax(smallc, sqrt(gamma * p_r / rho_r));
            double pstar = (W_l * p_r + W_r * p_l + W_l * W_r * (un_l - un_r)) / (W_l + W_r);
            pstar = pymax(pstar, smallp);
            const double ustar = (W_l * un_l + W_r * un_r + (p_l - p_r)) / (W_l + W_r);
            const double rhostar_l = rho_l + (pstar - p_l) / (c_l * c_l);
            const double rhostar_r = rho_r + (pstar - p_r) / (c_r * c_r);
            const double rhoestar_l = rhoe_l + (pstar - p_l) * (rhoe_l / rho_l + p_l / rho_l) / (c_l * c_l);
            const double rhoestar_r = rhoe_r + (pstar - p_r) * (rhoe_r / rho_r + p_r / rho_r) / (c_r * c_r);
            const double cstar_l = pymax(smallc, sqrt(gamma * pstar / rhostar_l));
            const double cstar_r = pymax(smallc, sqrt(gamma * pstar / rhostar_r));
            double rho_s, un_s, ut_s, rhoe_s;
            if (ustar > 0.0) {
                ut_s = ut_l;
                const double lam = un_l - c_l, lams = ustar - cstar_l;
                int star;          /* 1: star state, 0: left state, 2: inside the rarefaction */
                if (pstar > p_l) star = ((lam + lams) / 2.0 > 0.0) ? 0 : 1;
                else star = (lam < 0.0 && lams < 0.0) ? 1 : ((lam > 0.0 && lams > 0.0) ? 0 : 2);
                if (star == 0) { rho_s = rho_l; un_s = un_l; rhoe_s = rhoe_l; }
                else if (star == 1) { rho_s = rhostar_l; un_s = ustar; rhoe_s = rhoestar_l; }
                else {
                    const double alpha = lam / (lam - lams);
                    rho_s = alpha * rhostar_l + (1.0 - alpha) * rho_l;
                    un_s = alpha * ustar + (1.0 - alpha) * un_l;
                    rhoe_s = alpha * rhoestar_l + (1.0 - alpha) * rhoe_l;
                }
            } else if (ustar < 0) {
                ut_s = ut_r;
                const double lam = un_r + c_r, lams = ustar + cstar_r;
                int star;          /* 1: star state, 0: right state, 2: inside the rarefaction */
                if (pstar > p_r) star = ((lam + lams) / 2.0 > 0.0) ? 1 : 0;
                else star = (lam < 0.0 && lams < 0.0) ? 0 : ((lam > 0.0 && lams > 0.0) ? 1 : 2);
                if (star == 0) { rho_s = rho_r; un_s = un_r; rhoe_s = rhoe_r; }
                else if (star == 1) { rho_s = rhostar_r; un_s = ustar; rhoe_s = rhoestar_r; }
                else {
                    const double alpha = lam / (lam - lams);
                    rho_s = alpha * rhostar_r + (1.0 - alpha) * rho_r;
                    un_s = alpha * ustar + (1.0 - alpha) * un_r;
                    rhoe_s = alpha * rhoestar_r + (1.0 - alpha) * rhoe_r;
                }
            } else {
                rho_s = 0.5 * (rhostar_l + rhostar_r);
                un_s = ustar;
                ut_s = 0.5 * (ut_l + ut_r);
                rhoe_s = 0.5 * (rhoestar_l + rhoestar_r);
            }
            if (lower_solid && ((idir == 1 && i == ilo) || (idir == 2 && j == jlo))) un_s = 0.0;
            double Us[4], Fk[4];
            Us[IDENS] = rho_s;
            Us[imn] = rho_s * un_s;
            Us[imt] = rho_s * ut_s;
            Us[IENER] = rhoe_s + 0.5 * rho_s * (un_s * un_s + ut_s * ut_s);
            cons_flux_geom(idir, gamma, Us, Fk, coord_type);
            for (int m = 0; m < 4; m++) F[m * np + k] = Fk[m];
            if (Ustate)
                for (int m = 0; m < 4; m++) Ustate[m * np + k] = Us[m];
        }
}

/* riemann.py:864-1019 (riemann_hllc_lowspeed): HLLC in Toro's alternate form (Eqs. 10.43, 10.44) with the low Mach
 * number fix of Minoshima & Miyoshi (2021) -- the star-region pressure is blended towards the arithmetic mean as the
 * local Mach number drops */
static void riemann_hllc_lowspeed(int idir, const double *U_l, const double *U_r, double *F, int qx, int qy,
                                  int ng, double gamma)
{
    const size_t np = (size_t)qx * qy;
    const int nx = qx - 2 * ng, ny = qy - 2 * ng;
    const int ilo = ng, ihi = ng + nx, jlo = ng, jhi = ng + ny;
    const double smallc = 1.e-10, smallp = 1.e-10;
    const int imn = idir == 1 ? IXMOM : IYMOM, imt = idir == 1 ? IYMOM : IXMOM;
    memset(F, 0, 4 * np * sizeof(double));
#pragma omp parallel for
    for (int i = ilo - 1; i < ihi + 1; i++)
        for (int j = jlo - 1; j < jhi + 1; j++) {
            size_t k = IDX(i, j);
            double Ul[4], Ur[4], Fl[4], Fr[4], Fk[4], D[4] = {0.0, 0.0, 0.0, 0.0};
            for (int m = 0; m < 4; m++) { Ul[m] = U_l[m * np + k]; Ur[m] = U_r[m * np + k]; }
            double rho_l = Ul[IDENS];
            double un_l = Ul[imn] / rho_l, ut_l = Ul[imt] / rho_l;
            double rhoe_l = Ul[IENER] - 0.5 * rho_l * (un_l * un_l + ut_l * ut_l);
            double p_l = pymax(rhoe_l * (gamma - 1.0), smallp);
            double rho_r = Ur[IDENS];
            double un_r = Ur[imn] / rho_r, ut_r = Ur[imt] / rho_r;
            double rhoe_r = Ur[IENER] - 0.5 * rho_r * (un_r * un_r + ut_r * ut_r);
            double p_r = pymax(rhoe_r * (gamma - 1.0), smallp);
            double c_l = pymax(smallc, sqrt(gamma * p_l / rho_l));
            double c_r = pymax(smallc, sqrt(gamma * p_r / rho_r));
            double S_l, S_r;
            estimate_wave_speed(rho_l, un_l, p_l, c_l, rho_r, un_r, p_r, c_r, gamma, &S_l, &S_r);
            double S_c = (p_r - p_l + rho_l * un_l * (S_l - un_l) - rho_r * un_r * (S_r - un_r)) /
                         (rho_l * (S_l - un_l) - rho_r * (S_r - un_r));
            D[imn] = 1.0;
            D[IENER] = S_c;
            cons_flux(idir, gamma, Ul, Fl);
            cons_flux(idir, gamma, Ur, Fr);
            double vmag_l = sqrt(un_l * un_l + ut_l * ut_l);
            double vmag_r = sqrt(un_r * un_r + ut_r * ut_r);
            double cs_max = pymax(c_l, c_r);
            double chi = pymin(1.0, pymax(vmag_l, vmag_r) / cs_max);
            double phi = chi * (2.0 - chi);
            double pstar_lr = 0.5 * (p_l + p_r) +
                              0.5 * phi * (rho_l * (S_l - un_l) * (S_c - un_l) + rho_r * (S_r - un_r) * (S_c - un_r));
            if (S_r <= 0.0) {
                for (int m = 0; m < 4; m++) Fk[m] = Fr[m];
            } else if (S_c <= 0.0 && 0.0 < S_r) {
                for (int m = 0; m < 4; m++)
                    Fk[m] = (S_c * (S_r * Ur[m] - Fr[m]) + S_r * pstar_lr * D[m]) / (S_r - S_c);
            } else if (S_l < 0.0 && 0.0 < S_c) {
                for (int m = 0; m < 4; m++)
                    Fk[m] = (S_c * (S_l * Ul[m] - Fl[m]) + S_l * pstar_lr * D[m]) / (S_l - S_c);
            } else {
                for (int m = 0; m < 4; m++) Fk[m] = Fl[m];
            }
            for (int m = 0; m < 4; m++) F[m * np + k] = Fk[m];
        }
}

static void riemann_solve(int idir, const double *U_l, const double *U_r, double *F, int qx, int qy, int ng,
                          const orc_comp_params *P)
{
    if (P->riemann == 1) riemann_cgf(idir, U_l, U_r, F, qx, qy, ng, P->gamma, idir == 1 ? P->xl_solid : P->yl_solid);
    else if (P->riemann == 2) riemann_hllc_lowspeed(idir, U_l, U_r, F, qx, qy, ng, P->gamma);
    else riemann_hllc(idir, U_l, U_r, F, qx, qy, ng, P->gamma);
}

/* interface.py:240-378, Cartesian branch.  u, v full planes. */
static void artificial_viscosity(const double *u, const double *v, double *ax, double *ay, int qx,
                                 int qy, int ng, double dx, double dy, double cvisc, int skip_xhi,
                                 int skip_yhi, const orc_geom *G)
{
    const size_t np = (size_t)qx * qy;
    const int nx = qx - 2 * ng, ny = qy - 2 * ng;
    const int ilo = ng, ihi = ng + nx, jlo = ng, jhi = ng + ny;
    double *divU = zalloc(np);
    memset(ax, 0, np * sizeof(double));
    memset(ay, 0, np * sizeof(double));
#pragma omp parallel for
    for (int i = ilo - 1; i < ihi + 1; i++)
        for (int j = jlo - 1; j < jhi + 1; j++) {
            double ur = 0.5 * (u[IDX(i, j)] + u[IDX(i, j - 1)]);
            double ul = 0.5 * (u[IDX(i - 1, j)] + u[IDX(i - 1, j - 1)]);
            double vt = 0.5 * (v[IDX(i, j)] + v[IDX(i - 1, j)]);
            double vb = 0.5 * (v[IDX(i, j - 1)] + v[IDX(i - 1, j - 1)]);
            if (G) {
                /* divergence at the vertex in spherical polar coordinates (interface.py:332-353) */
                const double rr = (i + 0.5 - ng) * dx + G->xmin, rl = (i - 0.5 - ng) * dx + G->xmin;
                const double rc = (i - ng) * dx + G->xmin;
                const double ux = (ur * rr * rr - ul * rl * rl) / (rc * rc * dx);
                const double sint = sin((j + 0.5 - ng) * dy + G->ymin), sinb = sin((j - 0.5 - ng) * dy + G->ymin);
                const double sinc = sin((j - ng) * dy + G->ymin);
                const double vy = sinc == 0.0 ? 0.0 : (sint * vt - sinb * vb) / (rc * sinc * dy);
                divU[IDX(i, j)] = ux + vy;
                continue;
            }
            divU[IDX(i, j)] = (ur - ul) / dx + (vt - vb) / dy;
        }
    /* the reference loops range(ilo, ihi) x range(jlo, jhi): the +x / +y boundary faces are never
       set.  skip_*hi = 0 extends the loop by one face (used for interior slab
       boundaries in decomposed runs). */
    const int iend = skip_xhi ? ihi : ihi + 1, jend = skip_yhi ? jhi : jhi + 1;
#pragma omp parallel for
    for (int i = ilo; i < iend; i++)
        for (int j = jlo; j < jend; j++) {
            double divU_x = 0.5 * (divU[IDX(i, j)] + divU[IDX(i, j + 1)]);
            double divU_y = 0.5 * (divU[IDX(i, j)] + divU[IDX(i + 1, j)]);
            ax[IDX(i, j)] = cvisc * pymax(-divU_x * dx, 0.0);
            ay[IDX(i, j)] = cvisc * pymax(-divU_y * (G ? G->Ly[IDX(i, j)] : dy), 0.0);
        }
    free(divU);
}

static void dump(double *dst, const double *src, size_t n)
{
    if (dst) memcpy(dst, src, n * sizeof(double));
}

/* the compressible solver's "hse" boundary (compressible/BC.py:21-139) for ONE variable on ONE y side,
 * as fill_BC(name) applies it after the standard x fill: all variables but the energy copy the first
 * interior row; the energy integrates hydrostatic equilibrium outward at constant density from the
 * pressure of that row (over the whole x extent, whatever the other variables' x ghost cells hold at
 * that moment).  U = 4 planes (dens, ener, xmom, ymom); var = plane index; side 0 = ylb, 1 = yrb */
void orc_fill_hse(double *U, int nx, int ny, int ng, double dy, double grav, double gamma, int var, int side)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    const size_t np = (size_t)qx * qy;
    const int jb = side == 0 ? ng : ng + ny - 1;          /* jlo / jhi */
    const int step = side == 0 ? -1 : 1;
    double *v = U + (size_t)var * np;
    if (var != IENER) {
        for (int i = 0; i < qx; i++)
            for (int k = 1; k <= ng; k++) v[IDX(i, jb + step * k)] = v[IDX(i, jb)];
        return;
    }
    const double *dens = U + IDENS * np, *xmom = U + IXMOM * np, *ymom = U + IYMOM * np;
    for (int i = 0; i < qx; i++) {
        const size_t kb = IDX(i, jb);
        const double dens_base = dens[kb];
        const double ke_base = 0.5 * (xmom[kb] * xmom[kb] + ymom[kb] * ymom[kb]) / dens[kb];
        const double eint_base = (v[kb] - ke_base) / dens[kb];
        double pres_base = dens_base * eint_base * (gamma - 1.0);
        for (int k = 1; k <= ng; k++) {
            /* ylb: pres_below = pres_base - grav*dens_base*dy;  yrb: pres_above = pres_base + grav*dens_base*dy */
            const double pnext = side == 0 ? pres_base - grav * dens_base * dy : pres_base + grav * dens_base * dy;
            v[IDX(i, jb + step * k)] = pnext / (gamma - 1.0) + ke_base;
            pres_base = pnext;
        }
    }
}

/* one evolve() of compressible/simulation.py:290-450 (Cartesian, HLLC, gravity, no sponge,
 * no particles, no problem sources).  U (4 planes, ghosts already filled) is updated in place on the valid region.
 * returns 0, or 3 if the cons_to_prim assertion (simulation.py:71) would fire. */
int orc_compressible_step(double *U, int nx, int ny, int ng, double dx, double dy, double dt,
                          const orc_comp_params *P, const orc_comp_stages *S)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    const size_t np = (size_t)qx * qy;
    const double gamma = P->gamma;
    orc_comp_stages none;
    memset(&none, 0, sizeof none);
    if (!S) S = &none;
    const orc_geom *G = P->geom;       /* SphericalPolar: x = r, y = theta; CGF only (simulation.py:201-209) */
    if (G && P->riemann != 1) return 4;

    double *q = zalloc(4 * np), *xi = zalloc(np), *xi_x = zalloc(np), *xi_y = zalloc(np);
    double *ldx = zalloc(4 * np), *ldy = zalloc(4 * np), *tmp = zalloc(np), *tmp2 = zalloc(np);
    double *V_l = zalloc(4 * np), *V_r = zalloc(4 * np);
    double *U_xl = zalloc(4 * np), *U_xr = zalloc(4 * np), *U_yl = zalloc(4 * np), *U_yr = zalloc(4 * np);
    double *F_x = zalloc(4 * np), *F_y = zalloc(4 * np);
    int rc = 0;

    /* unsplit_fluxes.py:169 */
    if (cons_to_prim(U, q, qx, qy, ng, gamma)) rc = 3;
    dump(S->q, q, 4 * np);

    /* unsplit_fluxes.py:175-184 */
    if (P->use_flattening) {
        flatten1d(q, xi_x, qx, qy, ng, 1, P->z0, P->z1, P->delta);
        flatten1d(q, xi_y, qx, qy, ng, 2, P->z0, P->z1, P->delta);
        flatten_multid(q, xi_x, xi_y, xi, qx, qy, ng);
    } else {
        for (size_t k = 0; k < np; k++) xi[k] = 1.0;
    }
    dump(S->xi, xi, np);

    /* unsplit_fluxes.py:192-197 */
    for (int n = 0; n < 4; n++) {
        for (int idir = 1; idir <= 2; idir++) {
            double *out = (idir == 1 ? ldx : ldy) + n * np;
            if (P->limiter == 0) limit2(q + n * np, tmp2, qx, qy, ng, idir, 1);
            else if (P->limiter == 1) limit2(q + n * np, tmp2, qx, qy, ng, idir, 0);
            else limit4(q + n * np, tmp2, tmp, qx, qy, ng, idir);
            for (size_t k = 0; k < np; k++) out[k] = xi[k] * tmp2[k];
        }
    }
    dump(S->ldx, ldx, 4 * np);
    dump(S->ldy, ldy, 4 * np);

    /* unsplit_fluxes.py:207-242 */
    trace_states(1, q, ldx, V_l, V_r, qx, qy, ng, dx, dt, gamma, NULL, G ? G->dlogAx : NULL);
    prim_to_cons(V_l, U_xl, qx, qy, gamma);
    prim_to_cons(V_r, U_xr, qx, qy, gamma);
    trace_states(2, q, ldy, V_l, V_r, qx, qy, ng, dy, dt, gamma, G ? G->Ly : NULL, G ? G->dlogAy : NULL);
    prim_to_cons(V_l, U_yl, qx, qy, gamma);
    prim_to_cons(V_r, U_yr, qx, qy, gamma);
    dump(S->Uxl_hat, U_xl, 4 * np); dump(S->Uxr_hat, U_xr, 4 * np);
    dump(S->Uyl_hat, U_yl, 4 * np); dump(S->Uyr_hat, U_yr, 4 * np);

    /* apply_source_terms (unsplit_fluxes.py:247-330) with get_external_sources (simulation.py:105-128):
       S_ymom = dens * grav, S_ener = ymom * grav over the whole (ghost-filled) array, the source arrays
       then get THEIR OWN ghost fill, and half a time step of them goes to the buf = 1 interface states */
    if (P->grav != 0.0 || P->heat_profile || G) {
        double *src = zalloc(4 * np);
        for (size_t k = 0; k < np; k++) {
            if (G) {
                /* radial gravity plus the geometric (centrifugal / Coriolis-like) terms (simulation.py:117-124) */
                src[IXMOM * np + k] = U[IDENS * np + k] * P->grav;
                src[IENER * np + k] = U[IXMOM * np + k] * P->grav;
                src[IXMOM * np + k] += U[IYMOM * np + k] * U[IYMOM * np + k] / (U[IDENS * np + k] * G->x2d[k]);
                src[IYMOM * np + k] += -U[IXMOM * np + k] * U[IYMOM * np + k] / U[IDENS * np + k];
            } else {
                src[IYMOM * np + k] = U[IDENS * np + k] * P->grav;
                src[IENER * np + k] = U[IYMOM * np + k] * P->grav;
            }
            if (P->heat_profile) src[IENER * np + k] += U[IDENS * np + k] * P->heat_rate * P->heat_profile[k];
        }
        for (int n = 0; n < 4; n++)
            orc_fill_ghost_f64(src + n * np, nx, ny, ng, P->src_bc[4 * n], P->src_bc[4 * n + 1], P->src_bc[4 * n + 2],
                               P->src_bc[4 * n + 3], NULL, NULL, NULL, NULL, dx, dy);
        const int vars[3] = {IXMOM, IYMOM, IENER};
        for (int m = 0; m < 3; m++) {
            const int n = vars[m];
            const double *sv = src + n * np;
            double *xl = U_xl + n * np, *xr = U_xr + n * np, *yl = U_yl + n * np, *yr = U_yr + n * np;
            for (int i = ng - 1; i <= ng + nx; i++)
                for (int j = ng - 1; j <= ng + ny; j++) {
                    xl[IDX(i, j)] += 0.5 * dt * sv[IDX(i - 1, j)];
                    xr[IDX(i, j)] += 0.5 * dt * sv[IDX(i, j)];
                    yl[IDX(i, j)] += 0.5 * dt * sv[IDX(i, j - 1)];
                    yr[IDX(i, j)] += 0.5 * dt * sv[IDX(i, j)];
                }
        }
        free(src);
    }

    /* apply_transverse_flux (unsplit_fluxes.py:420-471) */
    double *Ust_x = NULL, *Ust_y = NULL, *qfx = NULL, *qfy = NULL;     /* SphericalPolar: interface states, their primitives */
    if (G) {
        Ust_x = zalloc(4 * np); Ust_y = zalloc(4 * np); qfx = zalloc(4 * np); qfy = zalloc(4 * np);
        riemann_cgf_geom(1, U_xl, U_xr, F_x, qx, qy, ng, gamma, P->xl_solid, 1, Ust_x);
        riemann_cgf_geom(2, U_yl, U_yr, F_y, qx, qy, ng, gamma, P->yl_solid, 1, Ust_y);
        cons_to_prim(Ust_x, qfx, qx, qy, ng, gamma);
        cons_to_prim(Ust_y, qfy, qx, qy, ng, gamma);
    } else {
        riemann_solve(1, U_xl, U_xr, F_x, qx, qy, ng, P);
        riemann_solve(2, U_yl, U_yr, F_y, qx, qy, ng, P);
    }
    dump(S->Fx_t, F_x, 4 * np); dump(S->Fy_t, F_y, 4 * np);
    if (G) {
        /* unsplit_fluxes.py:449-490 with face areas and cell volumes; note that the low-side states U_xl[i, j] /
           U_yl[i, j] use V, Ly, Lx of cell (i, j), the cell on the HIGH side of the face, like the reference */
        const double hdt = 0.5 * dt;
        const double *Axp = G->Ax, *Ayp = G->Ay, *pxf = qfx + IP * np, *pyf = qfy + IP * np;
        for (int n = 0; n < 4; n++) {
            double *xl = U_xl + n * np, *xr = U_xr + n * np, *yl = U_yl + n * np, *yr = U_yr + n * np;
            const double *fx = F_x + n * np, *fy = F_y + n * np;
            for (int i = ng - 2; i <= ng + nx; i++)
                for (int j = ng - 2; j <= ng + ny; j++) {
                    const double hdtV = hdt / G->V[IDX(i, j)];
                    xl[IDX(i, j)] += -hdtV * (fy[IDX(i - 1, j + 1)] * Ayp[IDX(i - 1, j + 1)] - fy[IDX(i - 1, j)] * Ayp[IDX(i - 1, j)]);
                    xr[IDX(i, j)] += -hdtV * (fy[IDX(i, j + 1)] * Ayp[IDX(i, j + 1)] - fy[IDX(i, j)] * Ayp[IDX(i, j)]);
                    yl[IDX(i, j)] += -hdtV * (fx[IDX(i + 1, j - 1)] * Axp[IDX(i + 1, j - 1)] - fx[IDX(i, j - 1)] * Axp[IDX(i, j - 1)]);
                    yr[IDX(i, j)] += -hdtV * (fx[IDX(i + 1, j)] * Axp[IDX(i + 1, j)] - fx[IDX(i, j)] * Axp[IDX(i, j)]);
                }
        }
        for (int i = ng - 2; i <= ng + nx; i++)
            for (int j = ng - 2; j <= ng + ny; j++) {
                const size_t k = IDX(i, j);
                U_xl[IYMOM * np + k] += -hdt * (pyf[IDX(i - 1, j + 1)] - pyf[IDX(i - 1, j)]) / G->Ly[k];
                U_xr[IYMOM * np + k] += -hdt * (pyf[IDX(i, j + 1)] - pyf[k]) / G->Ly[k];
                U_yl[IXMOM * np + k] += -hdt * (pxf[IDX(i + 1, j - 1)] - pxf[IDX(i, j - 1)]) / dx;
                U_yr[IXMOM * np + k] += -hdt * (pxf[IDX(i + 1, j)] - pxf[k]) / dx;
            }
    } else {
        const double hdt = 0.5 * dt, hdtV = hdt / (dx * dy), Ax = dy, Ay = dx;
        /* buf = (2, 1): i in [ilo-2, ihi+1], j likewise (inclusive ihi = ng+nx-1) */
        for (int n = 0; n < 4; n++) {
            double *xl = U_xl + n * np, *xr = U_xr + n * np, *yl = U_yl + n * np, *yr = U_yr + n * np;
            const double *fx = F_x + n * np, *fy = F_y + n * np;
#pragma omp parallel for
            for (int i = ng - 2; i <= ng + nx; i++)
                for (int j = ng - 2; j <= ng + ny; j++) {
                    xl[IDX(i, j)] += -hdtV * (fy[IDX(i - 1, j + 1)] * Ay - fy[IDX(i - 1, j)] * Ay);
                    xr[IDX(i, j)] += -hdtV * (fy[IDX(i, j + 1)] * Ay - fy[IDX(i, j)] * Ay);
                    yl[IDX(i, j)] += -hdtV * (fx[IDX(i + 1, j - 1)] * Ax - fx[IDX(i, j - 1)] * Ax);
                    yr[IDX(i, j)] += -hdtV * (fx[IDX(i + 1, j)] * Ax - fx[IDX(i, j)] * Ax);
                }
        }
    }
    dump(S->Uxl, U_xl, 4 * np); dump(S->Uxr, U_xr, 4 * np);
    dump(S->Uyl, U_yl, 4 * np); dump(S->Uyr, U_yr, 4 * np);

    /* final fluxes (simulation.py:330-357) */
    if (G) {
        memset(Ust_x, 0, 4 * np * sizeof(double)); memset(Ust_y, 0, 4 * np * sizeof(double));
        riemann_cgf_geom(1, U_xl, U_xr, F_x, qx, qy, ng, gamma, P->xl_solid, 1, Ust_x);
        riemann_cgf_geom(2, U_yl, U_yr, F_y, qx, qy, ng, gamma, P->yl_solid, 1, Ust_y);
        cons_to_prim(Ust_x, qfx, qx, qy, ng, gamma);
        cons_to_prim(Ust_y, qfy, qx, qy, ng, gamma);
    } else {
        riemann_solve(1, U_xl, U_xr, F_x, qx, qy, ng, P);
        riemann_solve(2, U_yl, U_yr, F_y, qx, qy, ng, P);
    }

    /* artificial viscosity (simulation.py:361-365, unsplit_fluxes.py:497-549) */
    if (cons_to_prim(U, q, qx, qy, ng, gamma)) rc = 3;
    {
        double *ax = tmp, *ay = tmp2;
        artificial_viscosity(q + IU * np, q + IV * np, ax, ay, qx, qy, ng, dx, dy, P->cvisc,
                             P->no_avisc_xhi, P->no_avisc_yhi, G);
        for (int n = 0; n < 4; n++) {
            const double *var = U + n * np;
            double *fx = F_x + n * np, *fy = F_y + n * np;
#pragma omp parallel for
            for (int i = ng - 2; i <= ng + nx; i++)
                for (int j = ng - 2; j <= ng + ny; j++) {
                    fx[IDX(i, j)] += ax[IDX(i, j)] * (var[IDX(i - 1, j)] - var[IDX(i, j)]);
                    fy[IDX(i, j)] += ay[IDX(i, j)] * (var[IDX(i, j - 1)] - var[IDX(i, j)]);
                }
        }
    }
    dump(S->Fx, F_x, 4 * np); dump(S->Fy, F_y, 4 * np);

    /* conservative update (simulation.py:377-384) */
    double *Uold_dens = zalloc(np), *Uold_ymom = zalloc(np), *Uold_xmom = zalloc(np);
    memcpy(Uold_dens, U + IDENS * np, np * sizeof(double));
    memcpy(Uold_ymom, U + IYMOM * np, np * sizeof(double));
    memcpy(Uold_xmom, U + IXMOM * np, np * sizeof(double));
    if (G) {
        /* simulation.py:377-396: flux differences with face areas over the cell volume, then the pressure gradients */
        for (int n = 0; n < 4; n++) {
            double *var = U + n * np;
            const double *fx = F_x + n * np, *fy = F_y + n * np;
            for (int i = ng; i < ng + nx; i++)
                for (int j = ng; j < ng + ny; j++) {
                    const double dtdV = dt / G->V[IDX(i, j)];
                    var[IDX(i, j)] += dtdV * (fx[IDX(i, j)] * G->Ax[IDX(i, j)] - fx[IDX(i + 1, j)] * G->Ax[IDX(i + 1, j)] +
                                              fy[IDX(i, j)] * G->Ay[IDX(i, j)] - fy[IDX(i, j + 1)] * G->Ay[IDX(i, j + 1)]);
                }
        }
        const double *pxf = qfx + IP * np, *pyf = qfy + IP * np;
        for (int i = ng; i < ng + nx; i++)
            for (int j = ng; j < ng + ny; j++) {
                U[IXMOM * np + IDX(i, j)] -= dt * (pxf[IDX(i + 1, j)] - pxf[IDX(i, j)]) / dx;
                U[IYMOM * np + IDX(i, j)] -= dt * (pyf[IDX(i, j + 1)] - pyf[IDX(i, j)]) / G->Ly[IDX(i, j)];
            }
    } else {
        const double dtdV = dt / (dx * dy), Ax = dy, Ay = dx;
        for (int n = 0; n < 4; n++) {
            double *var = U + n * np;
            const double *fx = F_x + n * np, *fy = F_y + n * np;
#pragma omp parallel for
            for (int i = ng; i < ng + nx; i++)
                for (int j = ng; j < ng + ny; j++)
                    var[IDX(i, j)] += dtdV * (fx[IDX(i, j)] * Ax - fx[IDX(i + 1, j)] * Ax +
                                              fy[IDX(i, j)] * Ay - fy[IDX(i, j + 1)] * Ay);
        }
    }
    /* external sources, predictor-corrector (simulation.py:398-423, get_external_sources :105-160):
       U += dt S(U_old); S_new uses the updated density and a time-centred y-momentum;
       U += dt/2 (S_new - S_old).  clean_state is a no-op for the default small_dens = -1e200 */
    if (G) {
        /* simulation.py:398-423 with the SphericalPolar branches of get_external_sources (:117-124, :135-146) */
        const double g = P->grav;
        for (int i = ng; i < ng + nx; i++)
            for (int j = ng; j < ng + ny; j++) {
                const size_t k = IDX(i, j);
                const double r = G->x2d[k];
                const double hp = P->heat_profile ? P->heat_profile[k] : 0.0;
                /* S_old from U_old */
                double so_x = Uold_dens[k] * g;
                double so_e = Uold_xmom[k] * g;
                if (P->heat_profile) so_e += Uold_dens[k] * P->heat_rate * hp;
                so_x += Uold_ymom[k] * Uold_ymom[k] / (Uold_dens[k] * r);
                const double so_y = 0.0 + -Uold_xmom[k] * Uold_ymom[k] / Uold_dens[k];
                U[IXMOM * np + k] += dt * so_x;
                U[IYMOM * np + k] += dt * so_y;
                U[IENER * np + k] += dt * so_e;
                /* S_new from the updated state, energy source with the time-centred radial momentum */
                double sn_x = U[IDENS * np + k] * g;
                const double so_xg = Uold_dens[k] * g;
                const double xmom_new = U[IXMOM * np + k] + 0.5 * dt * (sn_x - so_xg);
                double sn_e = xmom_new * g;
                if (P->heat_profile) sn_e += U[IDENS * np + k] * P->heat_rate * hp;
                sn_x += U[IYMOM * np + k] * U[IYMOM * np + k] / (U[IDENS * np + k] * r);
                const double sn_y = 0.0 + -U[IXMOM * np + k] * U[IYMOM * np + k] / U[IDENS * np + k];
                U[IXMOM * np + k] += 0.5 * dt * (sn_x - so_x);
                U[IYMOM * np + k] += 0.5 * dt * (sn_y - so_y);
                U[IENER * np + k] += 0.5 * dt * (sn_e - so_e);
            }
    } else if (P->grav != 0.0 || P->heat_profile) {
        const double g = P->grav;
#pragma omp parallel for
        for (int i = ng; i < ng + nx; i++)
            for (int j = ng; j < ng + ny; j++) {
                const size_t k = IDX(i, j);
                const double hp = P->heat_profile ? P->heat_profile[k] : 0.0;
                const double so_y = Uold_dens[k] * g;
                double so_e = Uold_ymom[k] * g;
                if (P->heat_profile) so_e += Uold_dens[k] * P->heat_rate * hp;
                U[IYMOM * np + k] += dt * so_y;
                U[IENER * np + k] += dt * so_e;
                const double sn_y = U[IDENS * np + k] * g;
                const double ymom_new = U[IYMOM * np + k] + 0.5 * dt * (sn_y - so_y);
                double sn_e = ymom_new * g;
                if (P->heat_profile) sn_e += U[IDENS * np + k] * P->heat_rate * hp;
                U[IYMOM * np + k] += 0.5 * dt * (sn_y - so_y);
                U[IENER * np + k] += 0.5 * dt * (sn_e - so_e);
            }
    }
    /* sponge: implicit damping of the momenta where the density is low, kinetic-energy change booked into the
       energy (simulation.py:425-441; the reference applies it to the whole array, ghost cells included, which the
       next fill overwrites) */
    if (P->do_sponge) {
        const double rb = P->sponge_rho_begin, rf = P->sponge_rho_full;
#pragma omp parallel for
        for (int i = ng; i < ng + nx; i++)
            for (int j = ng; j < ng + ny; j++) {
                const size_t k = IDX(i, j);
                const double rho = U[IDENS * np + k];
                const double f = rho > rb ? 0.0 : (rho < rf ? 1.0 : 0.5 * (1.0 - cos(3.14159265358979323846 * (rho - rb) / (rf - rb))));
                const double kappa = f / P->sponge_timescale;
                const double xo = U[IXMOM * np + k], yo = U[IYMOM * np + k];
                const double xn = xo / (1.0 + dt * kappa), yn = yo / (1.0 + dt * kappa);
                U[IXMOM * np + k] = xn;
                U[IYMOM * np + k] = yn;
                U[IENER * np + k] += 0.5 * ((xn * xn + yn * yn) - (xo * xo + yo * yo)) / rho;
            }
    }
    free(Uold_dens); free(Uold_ymom); free(Uold_xmom);
    free(Ust_x); free(Ust_y); free(qfx); free(qfy);

    free(q); free(xi); free(xi_x); free(xi_y); free(ldx); free(ldy); free(tmp); free(tmp2);
    free(V_l); free(V_r); free(U_xl); free(U_xr); free(U_yl); free(U_yr); free(F_x); free(F_y);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * multigrid (constant coefficient, ng = 1): (alpha - beta L) phi = f
 * a level is three planes v, f, r of (n+2)^2
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int nlevels;   /* level l has 2^(l+1) cells per side (MG.py:207) */
    int bc[4];     /* xl, xr, yl, yr codes; "dirichlet" = REFLECT_ODD, "neumann" = OUTFLOW */
    double alpha, beta;
    double xmin, xmax, ymin, ymax;
    int nsmooth, nsmooth_bottom;
    double **v, **f, **r; /* per level planes, owned */
    /* inhomogeneous values on the finest level only (MG.py:231-242); NULL = homogeneous */
    double *xl_val, *xr_val, *yl_val, *yr_val;
    /* variable-coefficient variant (variable_coeff_MG.py): per level the cell-centred coefficient and
       the edge coefficients eta_x, eta_y (already divided by dx^2, dy^2); NULL = constant coefficients */
    double **cc, **ex, **ey;
} orc_mg;

orc_mg *orc_mg_create(int nx, const int bc[4], double alpha, double beta, double xmin, double xmax,
                      double ymin, double ymax, int nsmooth, int nsmooth_bottom)
{
    orc_mg *m = (orc_mg *)calloc(1, sizeof(orc_mg));
    m->nlevels = (int)(log((double)nx) / log(2.0)); /* MG.py:197 */
    memcpy(m->bc, bc, sizeof m->bc);
    m->alpha = alpha; m->beta = beta;
    m->xmin = xmin; m->xmax = xmax; m->ymin = ymin; m->ymax = ymax;
    m->nsmooth = nsmooth; m->nsmooth_bottom = nsmooth_bottom;
    m->v = (double **)calloc(m->nlevels, sizeof(double *));
    m->f = (double **)calloc(m->nlevels, sizeof(double *));
    m->r = (double **)calloc(m->nlevels, sizeof(double *));
    int n = 2;
    for (int l = 0; l < m->nlevels; l++, n *= 2) {
        size_t np = (size_t)(n + 2) * (n + 2);
        m->v[l] = zalloc(np); m->f[l] = zalloc(np); m->r[l] = zalloc(np);
    }
    return m;
}

void orc_mg_destroy(orc_mg *m)
{
    for (int l = 0; l < m->nlevels; l++) { free(m->v[l]); free(m->f[l]); free(m->r[l]); }
    free(m->v); free(m->f); free(m->r);
    free(m->xl_val); free(m->xr_val); free(m->yl_val); free(m->yr_val);
    if (m->cc) {
        for (int l = 0; l < m->nlevels; l++) { free(m->cc[l]); free(m->ex[l]); free(m->ey[l]); }
        free(m->cc); free(m->ex); free(m->ey);
    }
    free(m);
}

int orc_mg_nlevels(const orc_mg *m) { return m->nlevels; }
double *orc_mg_plane(orc_mg *m, int level, int which)
{
    return which == 0 ? m->v[level] : which == 1 ? m->f[level] : m->r[level];
}

/* side: 0 xl, 1 xr, 2 yl, 3 yr; vals has n+2 entries for the finest level */
void orc_mg_set_bc_values(orc_mg *m, int side, const double *vals)
{
    int n = 1 << m->nlevels;
    double **dst = side == 0 ? &m->xl_val : side == 1 ? &m->xr_val : side == 2 ? &m->yl_val : &m->yr_val;
    free(*dst);
    *dst = NULL;
    if (vals) { *dst = zalloc(n + 2); memcpy(*dst, vals, (n + 2) * sizeof(double)); }
}

static int level_n(int l) { return 2 << l; }
static double level_dx(const orc_mg *m, int l) { return (m->xmax - m->xmin) / level_n(l); }

/* VarCoeffCCMG2d.__init__ (variable_coeff_MG.py:40-109) + EdgeCoeffs (edge_coeffs.py:1-54):
 * coeffs = finest-level cell-centred coefficient (n+2)^2 (valid cells used), cbc = its BC codes */
void orc_mg_set_coeffs(orc_mg *m, const double *coeffs, const int cbc[4])
{
    const int L = m->nlevels - 1;
    if (!m->cc) {
        m->cc = (double **)calloc(m->nlevels, sizeof(double *));
        m->ex = (double **)calloc(m->nlevels, sizeof(double *));
        m->ey = (double **)calloc(m->nlevels, sizeof(double *));
        for (int l = 0; l < m->nlevels; l++) {
            size_t np = (size_t)(level_n(l) + 2) * (level_n(l) + 2);
            m->cc[l] = zalloc(np); m->ex[l] = zalloc(np); m->ey[l] = zalloc(np);
        }
    }
    for (int l = L; l >= 0; l--) {
        const int n = level_n(l), qy = n + 2;
        const size_t np = (size_t)qy * qy;
        double *c = m->cc[l], *ex = m->ex[l], *ey = m->ey[l];
        const double dx = level_dx(m, l), dy = (m->ymax - m->ymin) / n;
        memset(c, 0, np * sizeof(double)); memset(ex, 0, np * sizeof(double)); memset(ey, 0, np * sizeof(double));
        if (l == L) {
            for (int i = 1; i <= n; i++)
                for (int j = 1; j <= n; j++) c[IDX(i, j)] = coeffs[IDX(i, j)];
        } else {
            /* coeffs_c.v() = f_patch.restrict("coeffs").v()  (patch.py:659-662) */
            const double *cf = m->cc[l + 1];
            const int qyf = level_n(l + 1) + 2;
            for (int i = 1; i <= n; i++)
                for (int j = 1; j <= n; j++) {
                    size_t k = (size_t)(2 * i - 1) * qyf + (2 * j - 1);
                    c[IDX(i, j)] = 0.25 * (cf[k] + cf[k + qyf] + cf[k + 1] + cf[k + qyf + 1]);
                }
        }
        orc_fill_ghost_f64(c, n, n, 1, cbc[0], cbc[1], cbc[2], cbc[3], NULL, NULL, NULL, NULL, dx, dy);
        if (l == L) {
            /* EdgeCoeffs(g, eta): region buf = (0, 1) */
            for (int i = 1; i <= n + 1; i++)
                for (int j = 1; j <= n + 1; j++) {
                    ex[IDX(i, j)] = 0.5 * (c[IDX(i - 1, j)] + c[IDX(i, j)]);
                    ey[IDX(i, j)] = 0.5 * (c[IDX(i, j - 1)] + c[IDX(i, j)]);
                }
            for (size_t k = 0; k < np; k++) { ex[k] /= dx * dx; ey[k] /= dy * dy; }
        } else {
            /* EdgeCoeffs.restrict() of the finer level's edge coefficients */
            const double *xf = m->ex[l + 1], *yf = m->ey[l + 1];
            const int qyf = level_n(l + 1) + 2;
            const double fdx = level_dx(m, l + 1), fdy = (m->ymax - m->ymin) / level_n(l + 1);
            for (int i = 1; i <= n + 1; i++)
                for (int j = 1; j <= n; j++) {
                    size_t k = (size_t)(2 * i - 1) * qyf + (2 * j - 1);
                    ex[IDX(i, j)] = 0.5 * (xf[k] + xf[k + 1]);
                }
            for (int i = 1; i <= n; i++)
                for (int j = 1; j <= n + 1; j++) {
                    size_t k = (size_t)(2 * i - 1) * qyf + (2 * j - 1);
                    ey[IDX(i, j)] = 0.5 * (yf[k] + yf[k + qyf]);
                }
            for (size_t k = 0; k < np; k++) {
                ex[k] = ex[k] * (fdx * fdx) / (dx * dx);
                ey[k] = ey[k] * (fdy * fdy) / (dy * dy);
            }
        }
    }
}

double *orc_mg_coef_plane(orc_mg *m, int level, int which)
{
    return which == 0 ? m->cc[level] : which == 1 ? m->ex[level] : m->ey[level];
}



static void mg_fill_bc_v(orc_mg *m, int l)
{
    int n = level_n(l), fin = (l == m->nlevels - 1);
    double dx = level_dx(m, l), dy = (m->ymax - m->ymin) / n;
    orc_fill_ghost_f64(m->v[l], n, n, 1, m->bc[0], m->bc[1], m->bc[2], m->bc[3],
                       fin ? m->xl_val : NULL, fin ? m->xr_val : NULL, fin ? m->yl_val : NULL,
                       fin ? m->yr_val : NULL, dx, dy);
}

/* MG.py:544-599 */
void orc_mg_smooth(orc_mg *m, int l, int nsmooth)
{
    const int n = level_n(l), qy = n + 2;
    double *v = m->v[l];
    const double *f = m->f[l];
    const double dx = level_dx(m, l), dy = (m->ymax - m->ymin) / n;
    const double xcoeff = m->beta / (dx * dx), ycoeff = m->beta / (dy * dy);
    static const int off[4][2] = {{0, 0}, {1, 1}, {1, 0}, {0, 1}};
    mg_fill_bc_v(m, l);
    if (m->ex) {
        /* variable_coeff_MG.py:137-171 */
        const double *ex = m->ex[l], *ey = m->ey[l];
        for (int it = 0; it < nsmooth; it++)
            for (int g = 0; g < 4; g++) {
                int ix = off[g][0], iy = off[g][1];
                for (int i = 1 + ix; i <= n; i += 2)
                    for (int j = 1 + iy; j <= n; j += 2) {
                        double denom = ex[IDX(i + 1, j)] + ex[IDX(i, j)] + ey[IDX(i, j + 1)] + ey[IDX(i, j)];
                        v[IDX(i, j)] = (-f[IDX(i, j)] + ex[IDX(i + 1, j)] * v[IDX(i + 1, j)] +
                                        ex[IDX(i, j)] * v[IDX(i - 1, j)] + ey[IDX(i, j + 1)] * v[IDX(i, j + 1)] +
                                        ey[IDX(i, j)] * v[IDX(i, j - 1)]) / denom;
                    }
                if (g == 1 || g == 3) mg_fill_bc_v(m, l);
            }
        return;
    }
    for (int it = 0; it < nsmooth; it++)
        for (int g = 0; g < 4; g++) {
            int ix = off[g][0], iy = off[g][1];
#pragma omp parallel for if (n >= 256)
            for (int i = 1 + ix; i <= n; i += 2)
                for (int j = 1 + iy; j <= n; j += 2)
                    v[IDX(i, j)] = (f[IDX(i, j)] + xcoeff * (v[IDX(i + 1, j)] + v[IDX(i - 1, j)]) +
                                    ycoeff * (v[IDX(i, j + 1)] + v[IDX(i, j - 1)])) /
                                   (m->alpha + 2.0 * xcoeff + 2.0 * ycoeff);
            if (g == 1 || g == 3) mg_fill_bc_v(m, l);
        }
}

/* MG.py:529-542 */
void orc_mg_residual(orc_mg *m, int l)
{
    const int n = level_n(l), qy = n + 2;
    const double *v = m->v[l], *f = m->f[l];
    double *r = m->r[l];
    const double dx = level_dx(m, l), dy = (m->ymax - m->ymin) / n;
    if (m->ex) {
        /* variable_coeff_MG.py:199-212: r = f - L_eta phi */
        const double *ex = m->ex[l], *ey = m->ey[l];
        for (int i = 1; i <= n; i++)
            for (int j = 1; j <= n; j++) {
                double L = ex[IDX(i + 1, j)] * (v[IDX(i + 1, j)] - v[IDX(i, j)]) -
                           ex[IDX(i, j)] * (v[IDX(i, j)] - v[IDX(i - 1, j)]) +
                           ey[IDX(i, j + 1)] * (v[IDX(i, j + 1)] - v[IDX(i, j)]) -
                           ey[IDX(i, j)] * (v[IDX(i, j)] - v[IDX(i, j - 1)]);
                r[IDX(i, j)] = f[IDX(i, j)] - L;
            }
        return;
    }
#pragma omp parallel for if (n >= 256)
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++)
            r[IDX(i, j)] = f[IDX(i, j)] - m->alpha * v[IDX(i, j)] +
                           m->beta * ((v[IDX(i - 1, j)] + v[IDX(i + 1, j)] - 2 * v[IDX(i, j)]) / (dx * dx) +
                                      (v[IDX(i, j - 1)] + v[IDX(i, j + 1)] - 2 * v[IDX(i, j)]) / (dy * dy));
}

/* patch.py:640-676 (N = 2): fine r -> coarse f, valid region only (MG.py:731-732) */
void orc_mg_restrict(orc_mg *m, int l)
{
    const int nc = level_n(l - 1), qyc = nc + 2, qy = level_n(l) + 2;
    const double *r = m->r[l];
    double *fc = m->f[l - 1];
#pragma omp parallel for if (nc >= 256)
    for (int ic = 1; ic <= nc; ic++)
        for (int jc = 1; jc <= nc; jc++) {
            int i = 2 * ic - 1, j = 2 * jc - 1;
            fc[(size_t)ic * qyc + jc] = 0.25 * (r[IDX(i, j)] + r[IDX(i + 1, j)] + r[IDX(i, j + 1)] + r[IDX(i + 1, j + 1)]);
        }
}

/* patch.py:678-736 + MG.py:745-751: v_fine += prolong(v_coarse); fill_BC */
void orc_mg_prolong_correct(orc_mg *m, int l)
{
    const int nc = level_n(l - 1), qyc = nc + 2, qy = level_n(l) + 2;
    const double *c = m->v[l - 1];
    double *v = m->v[l];
#define C(i, j) c[(size_t)(i) * qyc + (j)]
#pragma omp parallel for if (nc >= 256)
    for (int ic = 1; ic <= nc; ic++)
        for (int jc = 1; jc <= nc; jc++) {
            double mx = 0.5 * (C(ic + 1, jc) - C(ic - 1, jc));
            double my = 0.5 * (C(ic, jc + 1) - C(ic, jc - 1));
            double c0 = C(ic, jc);
            int i = 2 * ic - 1, j = 2 * jc - 1;
            v[IDX(i, j)] += c0 - 0.25 * mx - 0.25 * my;
            v[IDX(i + 1, j)] += c0 + 0.25 * mx - 0.25 * my;
            v[IDX(i, j + 1)] += c0 - 0.25 * mx + 0.25 * my;
            v[IDX(i + 1, j + 1)] += c0 + 0.25 * mx + 0.25 * my;
        }
#undef C
    mg_fill_bc_v(m, l);
}

/* MG.py:699-778 */
void orc_mg_vcycle(orc_mg *m, int l)
{
    if (l > 0) {
        orc_mg_smooth(m, l, m->nsmooth);
        orc_mg_residual(m, l);
        orc_mg_restrict(m, l);
        orc_mg_vcycle(m, l - 1);
        orc_mg_prolong_correct(m, l);
        orc_mg_smooth(m, l, m->nsmooth);
    } else {
        orc_mg_smooth(m, 0, m->nsmooth_bottom);
        mg_fill_bc_v(m, 0);
    }
}

/* array_indexer.py:98-111, valid region of an ng = 1 plane (plain left-to-right sum; numpy's
 * pairwise order differs at the 1e-16 level) */
double orc_norm(const double *a, int n, double dx, double dy)
{
    const int qy = n + 2;
    double s = 0.0;
    for (int i = 1; i <= n; i++) {
        double si = 0.0;
        for (int j = 1; j <= n; j++) si += a[IDX(i, j)] * a[IDX(i, j)];
        s += si;
    }
    return sqrt(dx * dy * s);
}

/* MG.py:623-697.  returns num_cycles; resid[c], relerr[c] get the per-cycle diagnostics */
int orc_mg_solve(orc_mg *m, double rtol, double source_norm, int max_cycles, double *resid,
                 double *relerr)
{
    const int L = m->nlevels - 1, n = level_n(L), qy = n + 2;
    const size_t np = (size_t)(n + 2) * (n + 2);
    const double dx = level_dx(m, L), dy = (m->ymax - m->ymin) / n;
    double *old_phi = zalloc(np), *diff = zalloc(np);
    memcpy(old_phi, m->v[L], np * sizeof(double));
    double residual_error = 1.e33;
    int cycle = 1;
    while (residual_error > rtol && cycle <= max_cycles) {
        for (int l = 0; l < L; l++) memset(m->v[l], 0, (size_t)(level_n(l) + 2) * (level_n(l) + 2) * sizeof(double));
        orc_mg_vcycle(m, L);
        for (int i = 1; i <= n; i++)
            for (int j = 1; j <= n; j++)
                diff[IDX(i, j)] = (m->v[L][IDX(i, j)] - old_phi[IDX(i, j)]) / (m->v[L][IDX(i, j)] + 1.e-16);
        double relative_error = orc_norm(diff, n, dx, dy);
        memcpy(old_phi, m->v[L], np * sizeof(double));
        orc_mg_residual(m, L);
        double rn = orc_norm(m->r[L], n, dx, dy);
        residual_error = source_norm != 0.0 ? rn / source_norm : rn;
        if (resid) resid[cycle - 1] = residual_error;
        if (relerr) relerr[cycle - 1] = relative_error;
        cycle++;
    }
    mg_fill_bc_v(m, L);
    free(old_phi); free(diff);
    return cycle - 1;
}

/* ------------------------------------------------------------------------------------------
 * Burgers / incompressible explicit part + the incompressible evolve() (uses the helpers above)
 * ---------------------------------------------------------------------------------------- */
#include "incomp_oracle.c"
#include "lm_oracle.c"
