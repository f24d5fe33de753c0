/* lm_oracle.c -- CPU restatement of pyro2's low Mach number atmospheric solver (lm_atm): the numba interface
 * routines and Simulation.evolve / preevolve / method_compute_timestep, with the variable-coefficient
 * multigrid oracle above doing the projections.  TEST INFRASTRUCTURE ONLY (#included by pyro_oracle.c).
 *
 * Reference (pyro2, file:line):
 *   LM_atm_interface.mac_vels / states            pyro/lm_atm/LM_atm_interface.py:181-329
 *   LM_atm_interface.rho_states                                                     :332-426
 *   LM_atm_interface.get_interface_states                                           :429-585
 *   LM_atm_interface.upwind / riemann / riemann_and_upwind                          :588-703
 *   Simulation.method_compute_timestep            pyro/lm_atm/simulation.py:138-178
 *   Simulation.preevolve                                                            :180-284
 *   Simulation.evolve                                                               :286-618
 *
 * The index ranges differ from the incompressible solver's (the Riemann / upwind loops run over
 * [ilo-1, ihi+1] with ihi = ng + nx, the transverse loop over [ilo-1, ihi]) and values outside them stay zero;
 * cell updates next to the boundary do read such partially-built entries, so the ranges are reproduced
 * exactly.  State = 8 planes of (n + 2 ng)^2: density, x-velocity, y-velocity, eint, phi-MAC, phi, gradp_x,
 * gradp_y; base state = 4 arrays of n + 2 ng: rho0, p0, beta0, beta0-edges.
 */

/* upwind / riemann over i in [ng-1, ng+nx+1], j likewise (LM_atm_interface.py:588-677) */
#define FOR_LM_RU(i, j) \
    for (int i = ng - 1; i <= ng + nx + 1; i++) \
        for (int j = ng - 1; j <= ng + ny + 1; j++)

static void lm_upwind(const double *ql, const double *qr, const double *s, double *q, int nx, int ny, int ng)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    memset(q, 0, (size_t)qx * qy * sizeof(double));
    FOR_LM_RU(i, j) {
        const size_t k = IDX(i, j);
        q[k] = s[k] > 0.0 ? ql[k] : (s[k] == 0.0 ? 0.5 * (ql[k] + qr[k]) : qr[k]);
    }
}

static void lm_riemann(const double *ql, const double *qr, double *s, int nx, int ny, int ng)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    memset(s, 0, (size_t)qx * qy * sizeof(double));
    FOR_LM_RU(i, j) {
        const size_t k = IDX(i, j);
        const double l = ql[k], r = qr[k];
        s[k] = (l > 0.0 && l + r > 0.0) ? l : ((l <= 0.0 && r >= 0.0) ? 0.0 : r);
    }
}

/* get_interface_states (LM_atm_interface.py:429-585): gpx, gpy already carry beta0/rho; source acts on v */
static faces8 lm_faces(const double *u, const double *v, const double *ldux, const double *ldvx, const double *lduy,
                       const double *ldvy, const double *gpx, const double *gpy, const double *source, int nx,
                       int ny, int ng, double dx, double dy, double dt)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    const size_t np = (size_t)qx * qy;
    const double dtdx = dt / dx, dtdy = dt / dy;
    faces8 S = faces_alloc(np);
    for (int i = ng - 2; i < ng + nx + 2; i++)
        for (int j = ng - 2; j < ng + ny + 2; j++) {
            const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
            const double uu = u[k], vv = v[k];
            S.u_xl[kx] = uu + 0.5 * (1.0 - dtdx * uu) * ldux[k];
            S.u_xr[k] = uu - 0.5 * (1.0 + dtdx * uu) * ldux[k];
            S.v_xl[kx] = vv + 0.5 * (1.0 - dtdx * uu) * ldvx[k];
            S.v_xr[k] = vv - 0.5 * (1.0 + dtdx * uu) * ldvx[k];
            S.u_yl[ky] = uu + 0.5 * (1.0 - dtdy * vv) * lduy[k];
            S.u_yr[k] = uu - 0.5 * (1.0 + dtdy * vv) * lduy[k];
            S.v_yl[ky] = vv + 0.5 * (1.0 - dtdy * vv) * ldvy[k];
            S.v_yr[k] = vv - 0.5 * (1.0 + dtdy * vv) * ldvy[k];
        }
    double *uhat = zalloc(np), *vhat = zalloc(np), *uxi = zalloc(np), *vxi = zalloc(np), *uyi = zalloc(np),
           *vyi = zalloc(np);
    lm_riemann(S.u_xl, S.u_xr, uhat, nx, ny, ng);
    lm_riemann(S.v_yl, S.v_yr, vhat, nx, ny, ng);
    lm_upwind(S.u_xl, S.u_xr, uhat, uxi, nx, ny, ng);
    lm_upwind(S.v_xl, S.v_xr, uhat, vxi, nx, ny, ng);
    lm_upwind(S.u_yl, S.u_yr, vhat, uyi, nx, ny, ng);
    lm_upwind(S.v_yl, S.v_yr, vhat, vyi, nx, ny, ng);
    for (int i = ng - 1; i < ng + nx + 1; i++)
        for (int j = ng - 1; j < ng + ny + 1; j++) {
            const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
            const double ubar = 0.5 * (uhat[k] + uhat[kx]);
            const double vbar = 0.5 * (vhat[k] + vhat[ky]);
            const double vu_y = vbar * (uyi[ky] - uyi[k]);
            S.u_xl[kx] = S.u_xl[kx] - 0.5 * dtdy * vu_y - 0.5 * dt * gpx[k];
            S.u_xr[k] = S.u_xr[k] - 0.5 * dtdy * vu_y - 0.5 * dt * gpx[k];
            const double vv_y = vbar * (vyi[ky] - vyi[k]);
            S.v_xl[kx] = S.v_xl[kx] - 0.5 * dtdy * vv_y - 0.5 * dt * gpy[k] + 0.5 * dt * source[k];
            S.v_xr[k] = S.v_xr[k] - 0.5 * dtdy * vv_y - 0.5 * dt * gpy[k] + 0.5 * dt * source[k];
            const double uv_x = ubar * (vxi[kx] - vxi[k]);
            S.v_yl[ky] = S.v_yl[ky] - 0.5 * dtdx * uv_x - 0.5 * dt * gpy[k] + 0.5 * dt * source[k];
            S.v_yr[k] = S.v_yr[k] - 0.5 * dtdx * uv_x - 0.5 * dt * gpy[k] + 0.5 * dt * source[k];
            const double uu_x = ubar * (uxi[kx] - uxi[k]);
            S.u_yl[ky] = S.u_yl[ky] - 0.5 * dtdx * uu_x - 0.5 * dt * gpx[k];
            S.u_yr[k] = S.u_yr[k] - 0.5 * dtdx * uu_x - 0.5 * dt * gpx[k];
        }
    free(uhat); free(vhat); free(uxi); free(vxi); free(uyi); free(vyi);
    return S;
}

/* rho_states (LM_atm_interface.py:332-426) */
static void lm_rho_states(const double *rho, const double *um, const double *vm, const double *ldrx,
                          const double *ldry, double *rxi, double *ryi, int nx, int ny, int ng, double dx, double dy,
                          double dt)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    const size_t np = (size_t)qx * qy;
    const double dtdx = dt / dx, dtdy = dt / dy;
    double *xl = zalloc(np), *xr = zalloc(np), *yl = zalloc(np), *yr = zalloc(np);
    for (int i = ng - 2; i < ng + nx + 2; i++)
        for (int j = ng - 2; j < ng + ny + 2; j++) {
            const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
            xl[kx] = rho[k] + 0.5 * (1.0 - dtdx * um[kx]) * ldrx[k];
            xr[k] = rho[k] - 0.5 * (1.0 + dtdx * um[k]) * ldrx[k];
            yl[ky] = rho[k] + 0.5 * (1.0 - dtdy * vm[ky]) * ldry[k];
            yr[k] = rho[k] - 0.5 * (1.0 + dtdy * vm[k]) * ldry[k];
        }
    lm_upwind(xl, xr, um, rxi, nx, ny, ng);
    lm_upwind(yl, yr, vm, ryi, nx, ny, ng);
    for (int i = ng - 2; i < ng + nx + 2; i++)
        for (int j = ng - 2; j < ng + ny + 2; j++) {
            const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
            const double u_x = (um[kx] - um[k]) / dx;
            const double v_y = (vm[ky] - vm[k]) / dy;
            const double rhov_y = (ryi[ky] * vm[ky] - ryi[k] * vm[k]) / dy;
            xl[kx] = xl[kx] - 0.5 * dt * (rhov_y + rho[k] * u_x);
            xr[k] = xr[k] - 0.5 * dt * (rhov_y + rho[k] * u_x);
            const double rhou_x = (rxi[kx] * um[kx] - rxi[k] * um[k]) / dx;
            yl[ky] = yl[ky] - 0.5 * dt * (rhou_x + rho[k] * v_y);
            yr[k] = yr[k] - 0.5 * dt * (rhou_x + rho[k] * v_y);
        }
    lm_upwind(xl, xr, um, rxi, nx, ny, ng);
    lm_upwind(yl, yr, vm, ryi, nx, ny, ng);
    free(xl); free(xr); free(yl); free(yr);
}

typedef struct {
    int n, ng;
    double xmin, xmax, ymin, ymax;
    double grav, gamma;
    int limiter, proj_type;
    int bc_dens[4], bc_xvel[4], bc_yvel[4], bc_phi[4];   /* BC codes; bc_phi in multigrid terms */
} orc_lm_params;

static void lm_fill(double *a, const orc_lm_params *P, const int *bc)
{
    const double dx = (P->xmax - P->xmin) / P->n, dy = (P->ymax - P->ymin) / P->n;
    orc_fill_ghost_f64(a, P->n, P->n, P->ng, bc[0], bc[1], bc[2], bc[3], NULL, NULL, NULL, NULL, dx, dy);
}

/* coeff.v() = numer / rho.v() [denominator given]; coeff.v() *= b[j] or b[j]^2 -- two roundings, as written */
static void lm_coeff_valid(double *coeff, const double *den, double numer, const double *b, int squared, int n, int ng,
                           int buf)
{
    const int qy = n + 2 * ng;
    for (int i = ng - buf; i < ng + n + buf; i++)
        for (int j = ng - buf; j < ng + n + buf; j++) {
            const size_t k = IDX(i, j);
            const double c = numer / den[k];
            coeff[k] = c * (squared ? b[j] * b[j] : b[j]);
        }
}

/* a VarCoeffCCMG2d solve: coefficients from the valid cells of `coeff`, RHS and (optional) initial guess given
 * on the multigrid grid; returns the hierarchy (caller destroys) */
static orc_mg *lm_project(const orc_lm_params *P, const double *coeff, const double *rhs, const double *guess,
                          double rtol, int *cycles)
{
    const int n = P->n, ng = P->ng, qy = n + 2 * ng, qm = n + 2;
    const size_t npm = (size_t)qm * qm;
    orc_mg *m = orc_mg_create(n, P->bc_phi, 0.0, 0.0, P->xmin, P->xmax, P->ymin, P->ymax, 10, 50);
    double *c = zalloc(npm);
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++) c[(size_t)i * qm + j] = coeff[IDX(i + ng - 1, j + ng - 1)];
    orc_mg_set_coeffs(m, c, P->bc_dens);
    free(c);
    const int L = m->nlevels - 1;
    const double snorm = mg_set_rhs(m, rhs);
    if (guess) memcpy(m->v[L], guess, npm * sizeof(double));
    *cycles = orc_mg_solve(m, rtol, snorm, 100, NULL, NULL);
    return m;
}

/* Simulation.evolve (lm_atm/simulation.py:286-618).  S = the 8 state planes (ghost cells of all variables
 * filled by the driver), base = rho0, p0, beta0, beta0-edges (n + 2 ng each).  The reference's aux_data
 * ("coeff", "source_y") is rewritten -- valid cells, then every ghost cell by its BC fill -- before each use,
 * so it is scratch here.  cycles[0..1] = V-cycles of the two projections. */
void orc_lm_evolve(double *S, const double *base, const orc_lm_params *P, double dt, int *cycles)
{
    const int n = P->n, ng = P->ng, nx = n, ny = n, qx = n + 2 * ng, qy = qx, qm = n + 2;
    const size_t np = (size_t)qx * qy, npm = (size_t)qm * qm;
    const double dx = (P->xmax - P->xmin) / n, dy = (P->ymax - P->ymin) / n, g = P->grav;
    double *rho = S, *u = S + np, *v = S + 2 * np, *eint = S + 3 * np, *phi_mac = S + 4 * np, *phi = S + 5 * np,
           *gpx = S + 6 * np, *gpy = S + 7 * np;
    const double *rho0 = base, *p0 = base + qy, *beta0 = base + 2 * qy, *b0e = base + 3 * qy;
    double *coeff = zalloc(np), *source = zalloc(np);

    double *ldrx = zalloc(np), *ldux = zalloc(np), *ldvx = zalloc(np), *ldry = zalloc(np), *lduy = zalloc(np),
           *ldvy = zalloc(np), *tmp = zalloc(np);
    slopes(rho, ldrx, tmp, qx, qy, ng, 1, P->limiter);
    slopes(u, ldux, tmp, qx, qy, ng, 1, P->limiter);
    slopes(v, ldvx, tmp, qx, qy, ng, 1, P->limiter);
    slopes(rho, ldry, tmp, qx, qy, ng, 2, P->limiter);
    slopes(u, lduy, tmp, qx, qy, ng, 2, P->limiter);
    slopes(v, ldvy, tmp, qx, qy, ng, 2, P->limiter);

    /* :309-325 coeff = beta0 / rho (with its ghost fill), source = rho' g / rho */
    lm_coeff_valid(coeff, rho, 1.0, beta0, 0, n, ng, 0);
    lm_fill(coeff, P, P->bc_dens);
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j < ng + n; j++) source[IDX(i, j)] = (rho[IDX(i, j)] - rho0[j]) * g / rho[IDX(i, j)];
    lm_fill(source, P, P->bc_yvel);

    double *cgx = zalloc(np), *cgy = zalloc(np);
    for (size_t k = 0; k < np; k++) { cgx[k] = coeff[k] * gpx[k]; cgy[k] = coeff[k] * gpy[k]; }
    faces8 F = lm_faces(u, v, ldux, ldvx, lduy, ldvy, cgx, cgy, source, nx, ny, ng, dx, dy, dt);
    double *um = zalloc(np), *vm = zalloc(np), *s = zalloc(np);
    lm_riemann(F.u_xl, F.u_xr, s, nx, ny, ng);
    lm_upwind(F.u_xl, F.u_xr, s, um, nx, ny, ng);
    lm_riemann(F.v_yl, F.v_yr, s, nx, ny, ng);
    lm_upwind(F.v_yl, F.v_yr, s, vm, nx, ny, ng);
    faces_free(F);

    /* MAC projection (:345-384) */
    lm_coeff_valid(coeff, rho, 1.0, beta0, 1, n, ng, 1);
    double *div = zalloc(npm);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const int I = i + ng, J = j + ng;
            div[(size_t)(i + 1) * qm + j + 1] = beta0[J] * (um[IDX(I + 1, J)] - um[IDX(I, J)]) / dx +
                                                (b0e[J + 1] * vm[IDX(I, J + 1)] - b0e[J] * vm[IDX(I, J)]) / dy;
        }
    orc_mg *m = lm_project(P, coeff, div, NULL, 1.e-12, &cycles[0]);
    memset(phi_mac, 0, np * sizeof(double));
    mg_to_grid(m->v[m->nlevels - 1], phi_mac, n, ng);
    orc_mg_destroy(m);

    /* MAC velocity correction with the edge-centred beta0/rho (:386-408) */
    lm_coeff_valid(coeff, rho, 1.0, beta0, 0, n, ng, 0);
    lm_fill(coeff, P, P->bc_dens);
    double *cx = zalloc(np), *cy = zalloc(np);
    for (int i = ng - 3; i <= ng + n; i++)
        for (int j = ng; j < ng + n; j++) cx[IDX(i, j)] = 0.5 * (coeff[IDX(i - 1, j)] + coeff[IDX(i, j)]);
    for (int i = ng; i < ng + n; i++)
        for (int j = ng - 3; j <= ng + n; j++) cy[IDX(i, j)] = 0.5 * (coeff[IDX(i, j - 1)] + coeff[IDX(i, j)]);
    for (int i = ng; i <= ng + n; i++)
        for (int j = ng; j < ng + n; j++) um[IDX(i, j)] -= cx[IDX(i, j)] * (phi_mac[IDX(i, j)] - phi_mac[IDX(i - 1, j)]) / dx;
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j <= ng + n; j++) vm[IDX(i, j)] -= cy[IDX(i, j)] * (phi_mac[IDX(i, j)] - phi_mac[IDX(i, j - 1)]) / dy;
    free(cx); free(cy);

    /* density update (:410-432) */
    double *rxi = zalloc(np), *ryi = zalloc(np), *rho_old = zalloc(np);
    lm_rho_states(rho, um, vm, ldrx, ldry, rxi, ryi, nx, ny, ng, dx, dy, dt);
    memcpy(rho_old, rho, np * sizeof(double));
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j < ng + n; j++) {
            const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
            rho[k] -= dt * ((rxi[kx] * um[kx] - rxi[k] * um[k]) / dx + (ryi[ky] * vm[ky] - ryi[k] * vm[k]) / dy);
        }
    lm_fill(rho, P, P->bc_dens);
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j < ng + n; j++) eint[IDX(i, j)] = p0[j] / (P->gamma - 1.0) / rho[IDX(i, j)];
    free(rxi); free(ryi);

    /* interface states of u, v with the time-centred density in the pressure term (:434-462) */
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j < ng + n; j++) {
            const size_t k = IDX(i, j);
            const double c = 2.0 / (rho[k] + rho_old[k]);
            coeff[k] = c * beta0[j];
        }
    lm_fill(coeff, P, P->bc_dens);
    for (size_t k = 0; k < np; k++) { cgx[k] = coeff[k] * gpx[k]; cgy[k] = coeff[k] * gpy[k]; }
    F = lm_faces(u, v, ldux, ldvx, lduy, ldvy, cgx, cgy, source, nx, ny, ng, dx, dy, dt);
    double *uxi = zalloc(np), *vxi = zalloc(np), *uyi = zalloc(np), *vyi = zalloc(np);
    lm_upwind(F.u_xl, F.u_xr, um, uxi, nx, ny, ng);
    lm_upwind(F.v_xl, F.v_xr, um, vxi, nx, ny, ng);
    lm_upwind(F.u_yl, F.u_yr, vm, uyi, nx, ny, ng);
    lm_upwind(F.v_yl, F.v_yr, vm, vyi, nx, ny, ng);
    faces_free(F);

    /* provisional velocity (:464-501) */
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j < ng + n; j++) {
            const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
            const double ub = 0.5 * (um[k] + um[kx]), vb = 0.5 * (vm[k] + vm[ky]);
            const double ax = ub * (uxi[kx] - uxi[k]) / dx + vb * (uyi[ky] - uyi[k]) / dy;
            const double ay = ub * (vxi[kx] - vxi[k]) / dx + vb * (vyi[ky] - vyi[k]) / dy;
            if (P->proj_type == 1) {
                u[k] -= (dt * ax + dt * gpx[k]);
                v[k] -= (dt * ay + dt * gpy[k]);
            } else {
                u[k] -= dt * ax;
                v[k] -= dt * ay;
            }
        }
    for (size_t k = 0; k < np; k++) {
        const int j = (int)(k % qy);
        const double rho_half = 0.5 * (rho[k] + rho_old[k]);
        source[k] = (rho_half - rho0[j]) * g / rho_half;
    }
    lm_fill(source, P, P->bc_yvel);
    for (size_t k = 0; k < np; k++) v[k] += dt * source[k];
    lm_fill(u, P, P->bc_xvel);
    lm_fill(v, P, P->bc_yvel);
    free(uxi); free(vxi); free(uyi); free(vyi); free(um); free(vm); free(s); free(rho_old);

    /* final projection (:512-598) */
    double *cfin = zalloc(np);
    lm_coeff_valid(cfin, rho, 1.0, beta0, 1, n, ng, 0);
    memset(div, 0, npm * sizeof(double));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const int I = i + ng, J = j + ng;
            div[(size_t)(i + 1) * qm + j + 1] = 0.5 * beta0[J] * (u[IDX(I + 1, J)] - u[IDX(I - 1, J)]) / dx +
                                                0.5 * (beta0[J + 1] * v[IDX(I, J + 1)] - beta0[J - 1] * v[IDX(I, J - 1)]) / dy;
        }
    for (size_t k = 0; k < npm; k++) div[k] = div[k] / dt;
    double *guess = zalloc(npm);
    for (int i = 0; i < n + 2; i++)
        for (int j = 0; j < n + 2; j++) guess[(size_t)i * qm + j] = phi[IDX(i + ng - 1, j + ng - 1)];
    m = lm_project(P, cfin, div, guess, 1.e-12, &cycles[1]);
    free(cfin); free(guess);
    memset(phi, 0, np * sizeof(double));
    const double *mv = m->v[m->nlevels - 1];
    mg_to_grid(mv, phi, n, ng);
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++) {
            const size_t k = IDX(i + ng - 1, j + ng - 1);
            const double gx = 0.5 * (mv[(size_t)(i + 1) * qm + j] - mv[(size_t)(i - 1) * qm + j]) / dx;
            const double gy = 0.5 * (mv[(size_t)i * qm + j + 1] - mv[(size_t)i * qm + j - 1]) / dy;
            const double c = (1.0 / rho[k]) * beta0[j + ng - 1];
            u[k] -= dt * c * gx;
            v[k] -= dt * c * gy;
            if (P->proj_type == 1) { gpx[k] += gx; gpy[k] += gy; }
            else { gpx[k] = gx; gpy[k] = gy; }
        }
    orc_mg_destroy(m);
    lm_fill(u, P, P->bc_xvel);
    lm_fill(v, P, P->bc_yvel);
    lm_fill(gpx, P, P->bc_dens);
    lm_fill(gpy, P, P->bc_dens);
    free(ldrx); free(ldux); free(ldvx); free(ldry); free(lduy); free(ldvy); free(tmp); free(cgx); free(cgy); free(div);
    free(coeff); free(source);
}

/* the initial projection of preevolve (lm_atm/simulation.py:180-262): L_coeff phi = D(beta0 U), U -= (beta0/rho) G phi,
 * with coeff = beta0^2 / rho; density and velocities are ghost-filled first.  Returns the V-cycle count. */
int orc_lm_initial_projection(double *S, const double *base, const orc_lm_params *P)
{
    const int n = P->n, ng = P->ng, qx = n + 2 * ng, qy = qx, qm = n + 2;
    const size_t np = (size_t)qx * qy, npm = (size_t)qm * qm;
    const double dx = (P->xmax - P->xmin) / n, dy = (P->ymax - P->ymin) / n;
    double *rho = S, *u = S + np, *v = S + 2 * np, *phi = S + 5 * np;
    const double *beta0 = base + 2 * qy;
    lm_fill(rho, P, P->bc_dens);
    lm_fill(u, P, P->bc_xvel);
    lm_fill(v, P, P->bc_yvel);
    double *coeff = zalloc(np), *div = zalloc(npm);
    lm_coeff_valid(coeff, rho, 1.0, beta0, 1, n, ng, 0);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const int I = i + ng, J = j + ng;
            div[(size_t)(i + 1) * qm + j + 1] = 0.5 * beta0[J] * (u[IDX(I + 1, J)] - u[IDX(I - 1, J)]) / dx +
                                                0.5 * (beta0[J + 1] * v[IDX(I, J + 1)] - beta0[J - 1] * v[IDX(I, J - 1)]) / dy;
        }
    int cycles = 0;
    orc_mg *m = lm_project(P, coeff, div, NULL, 1.e-10, &cycles);
    memset(phi, 0, np * sizeof(double));
    const double *mv = m->v[m->nlevels - 1];
    mg_to_grid(mv, phi, n, ng);
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++) {
            const size_t k = IDX(i + ng - 1, j + ng - 1);
            const double gx = 0.5 * (mv[(size_t)(i + 1) * qm + j] - mv[(size_t)(i - 1) * qm + j]) / dx;
            const double gy = 0.5 * (mv[(size_t)i * qm + j + 1] - mv[(size_t)i * qm + j - 1]) / dy;
            const double c = (1.0 / rho[k]) * beta0[j + ng - 1];
            u[k] -= c * gx;
            v[k] -= c * gy;
        }
    orc_mg_destroy(m);
    lm_fill(u, P, P->bc_xvel);
    lm_fill(v, P, P->bc_yvel);
    free(coeff); free(div);
    return cycles;
}

/* method_compute_timestep (lm_atm/simulation.py:138-178): advective CFL limit over the valid cells and the
 * buoyancy limit sqrt(2 dx / max(|rho' g| / rho)) */
double orc_lm_timestep(const double *S, const double *base, const orc_lm_params *P, double cfl)
{
    const int n = P->n, ng = P->ng, qx = n + 2 * ng, qy = qx;
    const size_t np = (size_t)qx * qy;
    const double dx = (P->xmax - P->xmin) / n, dy = (P->ymax - P->ymin) / n;
    const double *rho = S, *u = S + np, *v = S + 2 * np, *rho0 = base;
    double uall = 0.0, vall = 0.0, uval = 0.0, vval = 0.0, fb = 0.0;
    for (size_t k = 0; k < np; k++) {
        if (fabs(u[k]) > uall) uall = fabs(u[k]);
        if (fabs(v[k]) > vall) vall = fabs(v[k]);
    }
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j < ng + n; j++) {
            const size_t k = IDX(i, j);
            if (fabs(u[k]) > uval) uval = fabs(u[k]);
            if (fabs(v[k]) > vval) vval = fabs(v[k]);
            const double f = fabs((rho[k] - rho0[j]) * P->grav) / rho[k];
            if (f > fb) fb = f;
        }
    double xtmp = 1.e33, ytmp = 1.e33;
    if (!(uall == 0)) xtmp = dx / uval;
    if (!(vall == 0)) ytmp = dy / vval;
    const double dt = cfl * (xtmp < ytmp ? xtmp : ytmp);
    const double dt_buoy = sqrt(2.0 * dx / fb);
    return dt < dt_buoy ? dt : dt_buoy;
}

#undef FOR_LM_RU
