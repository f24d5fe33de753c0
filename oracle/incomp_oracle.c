/* incomp_oracle.c -- CPU restatement of the explicit (non-multigrid) part of pyro2's Burgers and
 * incompressible solvers, plus the full incompressible evolve() built on the multigrid oracle above.
 * TEST INFRASTRUCTURE ONLY (see pyro_oracle.c, which #includes this file at its end).
 *
 * Reference (pyro2, file:line):
 *   reconstruction.limit                        pyro/mesh/reconstruction.py:11-120 (limit2 / limit4 above)
 *   burgers_interface.get_interface_states      pyro/burgers/burgers_interface.py:4-79
 *   burgers_interface.apply_transverse_corrections                             :82-157
 *   burgers_interface.construct_unsplit_fluxes                                 :160-225
 *   burgers_interface.upwind / riemann / riemann_and_upwind                    :228-312
 *   burgers Simulation.evolve                   pyro/burgers/simulation.py:66-131
 *   incomp_interface.mac_vels / states          pyro/incompressible/incomp_interface.py:4-158
 *   incomp_interface.apply_gradp_corrections                                   :161-211
 *   incompressible Simulation.evolve            pyro/incompressible/simulation.py:159-404
 *
 * Every stage is a loop over the same index range the reference's slice expression covers, on
 * zero-initialised full-size arrays (grid.scratch_array()), with the reference's operation order, so
 * each stage is bit-identical to the reference (checked in tests/test_oracle_vs_reference.py).
 */

typedef struct {
    double *u_xl, *u_xr, *u_yl, *u_yr, *v_xl, *v_xr, *v_yl, *v_yr;
} faces8;

static faces8 faces_alloc(size_t np)
{
    faces8 S;
    S.u_xl = zalloc(np); S.u_xr = zalloc(np); S.u_yl = zalloc(np); S.u_yr = zalloc(np);
    S.v_xl = zalloc(np); S.v_xr = zalloc(np); S.v_yl = zalloc(np); S.v_yr = zalloc(np);
    return S;
}

static void faces_free(faces8 S)
{
    free(S.u_xl); free(S.u_xr); free(S.u_yl); free(S.u_yr);
    free(S.v_xl); free(S.v_xr); free(S.v_yl); free(S.v_yr);
}

/* loop over a.v(buf=2) */
#define FOR_BUF2(i, j) \
    for (int i = ng - 2; i <= qx - ng + 1; i++) \
        for (int j = ng - 2; j <= qy - ng + 1; j++)

static void slopes(const double *a, double *lda, double *tmp, int qx, int qy, int ng, int idir, int limiter)
{
    if (limiter == 0) limit2(a, lda, qx, qy, ng, idir, 1);
    else if (limiter == 1) limit2(a, lda, qx, qy, ng, idir, 0);
    else limit4(a, lda, tmp, qx, qy, ng, idir);
}

/* burgers_interface.py:45-77 */
static void burgers_states(const double *u, const double *v, const double *ldux, const double *ldvx,
                           const double *lduy, const double *ldvy, faces8 S, int qx, int qy, int ng,
                           double dtdx, double dtdy)
{
    FOR_BUF2(i, j) {
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
}

/* burgers_interface.py:257-284 (Almgren, Bell & Szymczak 1996) */
static void burgers_riemann(const double *ql, const double *qr, double *s, int qx, int qy, int ng)
{
    memset(s, 0, (size_t)qx * qy * sizeof(double));
    FOR_BUF2(i, j) {
        const size_t k = IDX(i, j);
        const double l = ql[k], r = qr[k];
        s[k] = (l <= 0.0 && r >= 0.0) ? 0.0 : ((l > 0.0 && l + r > 0.0) ? l : r);
    }
}

/* burgers_interface.py:228-254 */
static void burgers_upwind(const double *ql, const double *qr, const double *s, double *q, int qx, int qy, int ng)
{
    memset(q, 0, (size_t)qx * qy * sizeof(double));
    FOR_BUF2(i, j) {
        const size_t k = IDX(i, j);
        q[k] = (s[k] == 0.0) ? 0.5 * (ql[k] + qr[k]) : (s[k] > 0.0 ? ql[k] : qr[k]);
    }
}

/* burgers_interface.py:108-157 */
static void burgers_transverse(faces8 S, int qx, int qy, int ng, double dtdx, double dtdy)
{
    const size_t np = (size_t)qx * qy;
    double *uhat = zalloc(np), *vhat = zalloc(np), *uxi = zalloc(np), *vxi = zalloc(np), *uyi = zalloc(np),
           *vyi = zalloc(np);
    burgers_riemann(S.u_xl, S.u_xr, uhat, qx, qy, ng);
    burgers_riemann(S.v_yl, S.v_yr, vhat, qx, qy, ng);
    burgers_upwind(S.u_xl, S.u_xr, uhat, uxi, qx, qy, ng);
    burgers_upwind(S.v_xl, S.v_xr, uhat, vxi, qx, qy, ng);
    burgers_upwind(S.u_yl, S.u_yr, vhat, uyi, qx, qy, ng);
    burgers_upwind(S.v_yl, S.v_yr, vhat, vyi, qx, qy, ng);
    FOR_BUF2(i, j) {
        const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
        const double ubar = 0.5 * (uhat[k] + uhat[kx]);
        const double vbar = 0.5 * (vhat[k] + vhat[ky]);
        const double tu_y = -0.5 * dtdy * vbar * (uyi[ky] - uyi[k]);
        const double tv_y = -0.5 * dtdy * vbar * (vyi[ky] - vyi[k]);
        const double tv_x = -0.5 * dtdx * ubar * (vxi[kx] - vxi[k]);
        const double tu_x = -0.5 * dtdx * ubar * (uxi[kx] - uxi[k]);
        S.u_xl[kx] += tu_y; S.u_xr[k] += tu_y;
        S.v_xl[kx] += tv_y; S.v_xr[k] += tv_y;
        S.v_yl[ky] += tv_x; S.v_yr[k] += tv_x;
        S.u_yl[ky] += tu_x; S.u_yr[k] += tu_x;
    }
    free(uhat); free(vhat); free(uxi); free(vxi); free(uyi); free(vyi);
}

/* incomp_interface.py:190-209 */
static void incomp_gradp(faces8 S, const double *gpx, const double *gpy, int qx, int qy, int ng, double dt)
{
    FOR_BUF2(i, j) {
        const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
        const double cx = -0.5 * dt * gpx[k], cy = -0.5 * dt * gpy[k];
        S.u_xl[kx] += cx; S.u_xr[k] += cx;
        S.v_xl[kx] += cy; S.v_xr[k] += cy;
        S.v_yl[ky] += cy; S.v_yr[k] += cy;
        S.u_yl[ky] += cx; S.u_yr[k] += cx;
    }
}

/* the eight corrected interface states shared by mac_vels and states (incomp_interface.py:38-62, 105-129) */
static faces8 incomp_faces(const double *u, const double *v, const double *gpx, const double *gpy, int qx,
                           int qy, int ng, double dx, double dy, double dt, int limiter)
{
    const size_t np = (size_t)qx * qy;
    double *ldux = zalloc(np), *ldvx = zalloc(np), *lduy = zalloc(np), *ldvy = zalloc(np), *tmp = zalloc(np);
    slopes(u, ldux, tmp, qx, qy, ng, 1, limiter);
    slopes(v, ldvx, tmp, qx, qy, ng, 1, limiter);
    slopes(u, lduy, tmp, qx, qy, ng, 2, limiter);
    slopes(v, ldvy, tmp, qx, qy, ng, 2, limiter);
    faces8 S = faces_alloc(np);
    const double dtdx = dt / dx, dtdy = dt / dy;
    burgers_states(u, v, ldux, ldvx, lduy, ldvy, S, qx, qy, ng, dtdx, dtdy);
    burgers_transverse(S, qx, qy, ng, dtdx, dtdy);
    if (gpx) incomp_gradp(S, gpx, gpy, qx, qy, ng, dt);
    free(ldux); free(ldvx); free(lduy); free(ldvy); free(tmp);
    return S;
}

/* incomp_interface.mac_vels: u, v, gradp planes (qx*qy, ghost cells filled) -> u_MAC, v_MAC */
void orc_incomp_mac_vels(const double *u, const double *v, const double *gpx, const double *gpy, int nx, int ny,
                         int ng, double dx, double dy, double dt, int limiter, double *u_mac, double *v_mac)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    const size_t np = (size_t)qx * qy;
    faces8 S = incomp_faces(u, v, gpx, gpy, qx, qy, ng, dx, dy, dt, limiter);
    double *s = zalloc(np);
    burgers_riemann(S.u_xl, S.u_xr, s, qx, qy, ng);
    burgers_upwind(S.u_xl, S.u_xr, s, u_mac, qx, qy, ng);
    burgers_riemann(S.v_yl, S.v_yr, s, qx, qy, ng);
    burgers_upwind(S.v_yl, S.v_yr, s, v_mac, qx, qy, ng);
    free(s);
    faces_free(S);
}

/* incomp_interface.states: upwind all four interface velocities with the (projected) MAC velocities */
void orc_incomp_states(const double *u, const double *v, const double *gpx, const double *gpy, int nx, int ny,
                       int ng, double dx, double dy, double dt, int limiter, const double *u_mac,
                       const double *v_mac, double *u_xint, double *v_xint, double *u_yint, double *v_yint)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    faces8 S = incomp_faces(u, v, gpx, gpy, qx, qy, ng, dx, dy, dt, limiter);
    burgers_upwind(S.u_xl, S.u_xr, u_mac, u_xint, qx, qy, ng);
    burgers_upwind(S.v_xl, S.v_xr, u_mac, v_xint, qx, qy, ng);
    burgers_upwind(S.u_yl, S.u_yr, v_mac, u_yint, qx, qy, ng);
    burgers_upwind(S.v_yl, S.v_yr, v_mac, v_yint, qx, qy, ng);
    faces_free(S);
}

/* burgers Simulation.evolve (burgers/simulation.py:66-131) on ghost-filled u, v; updates the valid cells */
void orc_burgers_evolve(double *u, double *v, int nx, int ny, int ng, double dx, double dy, double dt, int limiter)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    const size_t np = (size_t)qx * qy;
    const double dtdx = dt / dx, dtdy = dt / dy;
    faces8 S = incomp_faces(u, v, NULL, NULL, qx, qy, ng, dx, dy, dt, limiter);
    double *s = zalloc(np), *um = zalloc(np), *vm = zalloc(np), *q = zalloc(np);
    double *fux = zalloc(np), *fvx = zalloc(np), *fuy = zalloc(np), *fvy = zalloc(np);
    burgers_riemann(S.u_xl, S.u_xr, s, qx, qy, ng);
    burgers_upwind(S.u_xl, S.u_xr, s, um, qx, qy, ng);
    burgers_riemann(S.v_yl, S.v_yr, s, qx, qy, ng);
    burgers_upwind(S.v_yl, S.v_yr, s, vm, qx, qy, ng);
    /* burgers_interface.py:206-223: f = 0.5 * q_int * MAC velocity */
    burgers_upwind(S.u_xl, S.u_xr, um, q, qx, qy, ng);
    FOR_BUF2(i, j) fux[IDX(i, j)] = 0.5 * q[IDX(i, j)] * um[IDX(i, j)];
    burgers_upwind(S.v_xl, S.v_xr, um, q, qx, qy, ng);
    FOR_BUF2(i, j) fvx[IDX(i, j)] = 0.5 * q[IDX(i, j)] * um[IDX(i, j)];
    burgers_upwind(S.u_yl, S.u_yr, vm, q, qx, qy, ng);
    FOR_BUF2(i, j) fuy[IDX(i, j)] = 0.5 * q[IDX(i, j)] * vm[IDX(i, j)];
    burgers_upwind(S.v_yl, S.v_yr, vm, q, qx, qy, ng);
    FOR_BUF2(i, j) fvy[IDX(i, j)] = 0.5 * q[IDX(i, j)] * vm[IDX(i, j)];
    for (int i = ng; i < ng + nx; i++)
        for (int j = ng; j < ng + ny; j++) {
            const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
            u[k] = u[k] + dtdx * (fux[k] - fux[kx]) + dtdy * (fuy[k] - fuy[ky]);
            v[k] = v[k] + dtdx * (fvx[k] - fvx[kx]) + dtdy * (fvy[k] - fvy[ky]);
        }
    free(s); free(um); free(vm); free(q); free(fux); free(fvx); free(fuy); free(fvy);
    faces_free(S);
}

/* copy the (n+2)^2 multigrid plane into / out of the buf = 1 region of an ng-ghost solver array */
static void mg_to_grid(const double *mgp, double *a, int n, int ng)
{
    const int qy = n + 2 * ng, qm = n + 2;
    for (int i = 0; i < n + 2; i++)
        for (int j = 0; j < n + 2; j++) a[IDX(i + ng - 1, j + ng - 1)] = mgp[(size_t)i * qm + j];
}

static double mg_set_rhs(orc_mg *m, const double *f)
{
    const int L = m->nlevels - 1, n = level_n(L);
    memcpy(m->f[L], f, (size_t)(n + 2) * (n + 2) * sizeof(double));
    return orc_norm(m->f[L], n, level_dx(m, L), (m->ymax - m->ymin) / n);
}

/* incompressible Simulation.evolve (incompressible/simulation.py:159-404), no extra sources.
 * S = 6 planes of (n + 2 ng)^2: x-velocity, y-velocity, phi-MAC, phi, gradp_x, gradp_y, ghost cells of
 * the velocities filled (the driver's fill_BC_all).  vel_bc = the BC codes of u then v (4 + 4), phi_bc
 * those of phi.  cycles[0..1] = V-cycles taken by the MAC and the final projection.
 * dump (optional): u_MAC, v_MAC after the MAC projection and the four upwinded interface states. */
void orc_incomp_evolve(double *S, int n, int ng, double xmin, double xmax, double ymin, double ymax, double dt,
                       int limiter, int proj_type, const int *vel_bc, const int *phi_bc, int *cycles,
                       double *dump)
{
    const int qx = n + 2 * ng, qy = qx, qm = n + 2;
    const size_t np = (size_t)qx * qy, npm = (size_t)qm * qm;
    const double dx = (xmax - xmin) / n, dy = (ymax - ymin) / n;
    double *u = S, *v = S + np, *phi_mac = S + 2 * np, *phi = S + 3 * np, *gpx = S + 4 * np, *gpy = S + 5 * np;
    double *um = zalloc(np), *vm = zalloc(np);
    orc_incomp_mac_vels(u, v, gpx, gpy, n, n, ng, dx, dy, dt, limiter, um, vm);

    /* MAC projection (:243-284) */
    orc_mg *m = orc_mg_create(n, phi_bc, 0.0, -1.0, xmin, xmax, ymin, ymax, 10, 50);
    const int L = m->nlevels - 1;
    double *div = zalloc(npm);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const size_t k = IDX(i + ng, j + ng);
            div[(size_t)(i + 1) * qm + j + 1] = (um[IDX(i + ng + 1, j + ng)] - um[k]) / dx + (vm[IDX(i + ng, j + ng + 1)] - vm[k]) / dy;
        }
    memset(m->v[L], 0, npm * sizeof(double));
    double snorm = mg_set_rhs(m, div);
    cycles[0] = orc_mg_solve(m, 1.e-12, snorm, 100, NULL, NULL);
    mg_to_grid(m->v[L], phi_mac, n, ng);
    for (int i = ng; i <= ng + n; i++)
        for (int j = ng; j < ng + n; j++) um[IDX(i, j)] -= (phi_mac[IDX(i, j)] - phi_mac[IDX(i - 1, j)]) / dx;
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j <= ng + n; j++) vm[IDX(i, j)] -= (phi_mac[IDX(i, j)] - phi_mac[IDX(i, j - 1)]) / dy;

    /* interface states upwinded with the MAC velocities, advective update (:286-336) */
    double *uxi = zalloc(np), *vxi = zalloc(np), *uyi = zalloc(np), *vyi = zalloc(np);
    orc_incomp_states(u, v, gpx, gpy, n, n, ng, dx, dy, dt, limiter, um, vm, uxi, vxi, uyi, vyi);
    if (dump) {
        memcpy(dump, um, np * 8); memcpy(dump + np, vm, np * 8); memcpy(dump + 2 * np, uxi, np * 8);
        memcpy(dump + 3 * np, vxi, np * 8); memcpy(dump + 4 * np, uyi, np * 8); memcpy(dump + 5 * np, vyi, np * 8);
    }
    double *ax = zalloc(np), *ay = zalloc(np);
    for (int i = ng; i < ng + n; i++)
        for (int j = ng; j < ng + n; j++) {
            const size_t k = IDX(i, j), kx = IDX(i + 1, j), ky = IDX(i, j + 1);
            const double ub = 0.5 * (um[k] + um[kx]), vb = 0.5 * (vm[k] + vm[ky]);
            ax[k] = ub * (uxi[kx] - uxi[k]) / dx + vb * (uyi[ky] - uyi[k]) / dy;
            ay[k] = ub * (vxi[kx] - vxi[k]) / dx + vb * (vyi[ky] - vyi[k]) / dy;
        }
    for (size_t k = 0; k < np; k++) {
        if (proj_type == 1) {
            u[k] -= (dt * ax[k] + dt * gpx[k]);
            v[k] -= (dt * ay[k] + dt * gpy[k]);
        } else {
            u[k] -= dt * ax[k];
            v[k] -= dt * ay[k];
        }
    }
    orc_fill_ghost_f64(u, n, n, ng, vel_bc[0], vel_bc[1], vel_bc[2], vel_bc[3], NULL, NULL, NULL, NULL, dx, dy);
    orc_fill_ghost_f64(v, n, n, ng, vel_bc[4], vel_bc[5], vel_bc[6], vel_bc[7], NULL, NULL, NULL, NULL, dx, dy);

    /* final projection (:343-393): a fresh hierarchy, RHS = div(U) / dt, phi as the initial guess */
    orc_mg_destroy(m);
    m = orc_mg_create(n, phi_bc, 0.0, -1.0, xmin, xmax, ymin, ymax, 10, 50);
    memset(div, 0, npm * sizeof(double));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const int I = i + ng, J = j + ng;
            div[(size_t)(i + 1) * qm + j + 1] = 0.5 * (u[IDX(I + 1, J)] - u[IDX(I - 1, J)]) / dx + 0.5 * (v[IDX(I, J + 1)] - v[IDX(I, J - 1)]) / dy;
        }
    for (size_t k = 0; k < npm; k++) div[k] = div[k] / dt;
    snorm = mg_set_rhs(m, div);
    for (int i = 0; i < n + 2; i++)
        for (int j = 0; j < n + 2; j++) m->v[L][(size_t)i * qm + j] = phi[IDX(i + ng - 1, j + ng - 1)];
    cycles[1] = orc_mg_solve(m, 1.e-12, snorm, 100, NULL, NULL);
    memset(phi, 0, np * sizeof(double));
    mg_to_grid(m->v[L], phi, n, ng);
    /* get_solution_gradient (MG.py:437-466): valid cells of the solver grid, zero elsewhere */
    const double *mv = m->v[L];
    for (size_t k = 0; k < np; k++) { ax[k] = 0.0; ay[k] = 0.0; }
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++) {
            const size_t k = IDX(i + ng - 1, j + ng - 1);
            ax[k] = 0.5 * (mv[(size_t)(i + 1) * qm + j] - mv[(size_t)(i - 1) * qm + j]) / dx;
            ay[k] = 0.5 * (mv[(size_t)i * qm + j + 1] - mv[(size_t)i * qm + j - 1]) / dy;
        }
    for (size_t k = 0; k < np; k++) {
        u[k] -= dt * ax[k];
        v[k] -= dt * ay[k];
        if (proj_type == 1) { gpx[k] += ax[k]; gpy[k] += ay[k]; }
        else { gpx[k] = ax[k]; gpy[k] = ay[k]; }
    }
    orc_fill_ghost_f64(u, n, n, ng, vel_bc[0], vel_bc[1], vel_bc[2], vel_bc[3], NULL, NULL, NULL, NULL, dx, dy);
    orc_fill_ghost_f64(v, n, n, ng, vel_bc[4], vel_bc[5], vel_bc[6], vel_bc[7], NULL, NULL, NULL, NULL, dx, dy);
    orc_mg_destroy(m);
    free(um); free(vm); free(div); free(uxi); free(vxi); free(uyi); free(vyi); free(ax); free(ay);
}

/* linear advection: advection/interface.py:linear_interface + advective_fluxes.py:unsplit_fluxes +
 * advection/simulation.py:evolve (:56-92); a = one ghost-filled plane, updated in its valid cells */
void orc_advection_evolve(double *a, int nx, int ny, int ng, double dx, double dy, double dt, double u, double v,
                          int limiter)
{
    const int qx = nx + 2 * ng, qy = ny + 2 * ng;
    const size_t np = (size_t)qx * qy;
    double *ldx = zalloc(np), *ldy = zalloc(np), *tmp = zalloc(np), *ax = zalloc(np), *ay = zalloc(np);
    double *fxt = zalloc(np), *fyt = zalloc(np), *fx = zalloc(np), *fy = zalloc(np);
    slopes(a, ldx, tmp, qx, qy, ng, 1, limiter);
    slopes(a, ldy, tmp, qx, qy, ng, 2, limiter);
    const double cx = u * dt / dx, cy = v * dt / dy;
    for (int i = ng - 1; i <= ng + nx; i++)
        for (int j = ng - 1; j <= ng + ny; j++) {
            const size_t k = IDX(i, j);
            ax[k] = (u < 0) ? a[k] - 0.5 * (1.0 + cx) * ldx[k] : a[IDX(i - 1, j)] + 0.5 * (1.0 - cx) * ldx[IDX(i - 1, j)];
            ay[k] = (v < 0) ? a[k] - 0.5 * (1.0 + cy) * ldy[k] : a[IDX(i, j - 1)] + 0.5 * (1.0 - cy) * ldy[IDX(i, j - 1)];
        }
    for (size_t k = 0; k < np; k++) { fxt[k] = u * ax[k]; fyt[k] = v * ay[k]; }
    const int mx = (u <= 0) ? 0 : -1, my = (v <= 0) ? 0 : -1;
    const double dtdx2 = 0.5 * dt / dx, dtdy2 = 0.5 * dt / dy;
    for (int i = ng - 1; i <= ng + nx; i++)
        for (int j = ng - 1; j <= ng + ny; j++) {
            const size_t k = IDX(i, j);
            fx[k] = u * (ax[k] - dtdy2 * (fyt[IDX(i + mx, j + 1)] - fyt[IDX(i + mx, j)]));
            fy[k] = v * (ay[k] - dtdx2 * (fxt[IDX(i + 1, j + my)] - fxt[IDX(i, j + my)]));
        }
    const double dtdx = dt / dx, dtdy = dt / dy;
    for (int i = ng; i < ng + nx; i++)
        for (int j = ng; j < ng + ny; j++) {
            const size_t k = IDX(i, j);
            a[k] = a[k] + dtdx * (fx[k] - fx[IDX(i + 1, j)]) + dtdy * (fy[k] - fy[IDX(i, j + 1)]);
        }
    free(ldx); free(ldy); free(tmp); free(ax); free(ay); free(fxt); free(fyt); free(fx); free(fy);
}

/* diffusion Simulation.evolve (pyro/diffusion/simulation.py:62-104): Crank-Nicolson with a multigrid solve of
 * (1 - dt k/2 L) phi^{n+1} = phi^n + dt k/2 L phi^n.  phi = one (n+2)^2 plane (ng = 1), updated in its valid
 * cells; returns the V-cycle count */
int orc_diffusion_evolve(double *phi, int n, double xmin, double xmax, double ymin, double ymax, double dt, double k,
                         const int *bc)
{
    const int qy = n + 2;
    const size_t np = (size_t)qy * qy;
    const double dx = (xmax - xmin) / n, dy = (ymax - ymin) / n;
    orc_fill_ghost_f64(phi, n, n, 1, bc[0], bc[1], bc[2], bc[3], NULL, NULL, NULL, NULL, dx, dy);
    orc_mg *m = orc_mg_create(n, bc, 1.0, 0.5 * dt * k, xmin, xmax, ymin, ymax, 10, 50);
    const int L = m->nlevels - 1;
    double *f = zalloc(np);
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++) {
            const size_t c = IDX(i, j);
            f[c] = phi[c] + 0.5 * dt * k * ((phi[IDX(i + 1, j)] + phi[IDX(i - 1, j)] - 2.0 * phi[c]) / (dx * dx) +
                                            (phi[IDX(i, j + 1)] + phi[IDX(i, j - 1)] - 2.0 * phi[c]) / (dy * dy));
        }
    const double snorm = mg_set_rhs(m, f);
    memset(m->v[L], 0, np * sizeof(double));
    const int cycles = orc_mg_solve(m, 1.e-10, snorm, 100, NULL, NULL);
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++) phi[IDX(i, j)] = m->v[L][IDX(i, j)];
    orc_mg_destroy(m);
    free(f);
    return cycles;
}

#undef FOR_BUF2
