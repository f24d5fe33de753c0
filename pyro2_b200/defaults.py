"""Default runtime parameters, as data.

The reference keeps these in ``_defaults`` text files next to each solver (loaded at
pyro/pyro_sim.py:83-85); here they are plain dictionaries ``{"section.key": (value, comment)}`` fed
to ``RuntimeParameters.load_dict``.  User ``inputs`` files in the reference's ``[section] key = value``
syntax are still accepted by ``Pyro.initialize_problem(inputs_file=...)``.
"""

# driver / io / mesh parameters every solver understands
GLOBAL = {
    "driver.tmax": (1.0, "stop time"),
    "driver.max_steps": (10000, "stop after this many steps"),
    "driver.fix_dt": (-1.0, "fixed timestep when > 0"),
    "driver.init_tstep_factor": (0.01, "shrink the very first CFL timestep by this factor"),
    "driver.max_dt_change": (2.0, "largest allowed growth of dt from one step to the next"),
    "driver.verbose": (1.0, "chattiness"),
    "driver.cfl": (0.8, "advective CFL number"),
    "io.basename": ("pyro_", "prefix of output files"),
    "io.dt_out": (0.1, "simulation time between outputs"),
    "io.n_out": (10000, "steps between outputs"),
    "io.do_io": (1, "write output files at all?"),
    "io.force_final_output": (0, "write the last state even if do_io is off"),
    "vis.dovis": (0, "runtime visualisation (not provided by the device build)"),
    "vis.store_images": (0, ""),
    "mesh.grid_type": ("Cartesian2d", "Cartesian2d or SphericalPolar (x = r, y = theta; compressible solver with CGF)"),
    "mesh.xmin": (0.0, ""), "mesh.xmax": (1.0, ""), "mesh.ymin": (0.0, ""), "mesh.ymax": (1.0, ""),
    "mesh.xlboundary": ("reflect", "reflect, outflow or periodic"),
    "mesh.xrboundary": ("reflect", ""),
    "mesh.ylboundary": ("reflect", ""),
    "mesh.yrboundary": ("reflect", ""),
    "mesh.nx": (25, "zones in x"), "mesh.ny": (25, "zones in y"),
    "particles.do_particles": (0, "tracer particles are not supported"),
}

SOLVER = {
    "compressible": {
        "driver.cfl": (0.8, ""),
        "eos.gamma": (1.4, "p = rho e (gamma - 1)"),
        "compressible.use_flattening": (1, "flatten slopes at shocks"),
        "compressible.z0": (0.75, "flattening parameter"),
        "compressible.z1": (0.85, "flattening parameter"),
        "compressible.delta": (0.33, "flattening parameter"),
        "compressible.cvisc": (0.1, "artificial viscosity coefficient"),
        "compressible.limiter": (2, "0 none, 1 second-order MC, 2 fourth-order MC"),
        "compressible.grav": (0.0, "constant gravitational acceleration along y"),
        "compressible.riemann": ("HLLC", "HLLC or CGF"),
        "compressible.small_dens": (-1.e200, "density floor"),
        "compressible.small_eint": (-1.e200, "internal-energy floor"),
        "sponge.do_sponge": (0, "damp the velocities in the low-density region above the atmosphere"),
        "sponge.sponge_rho_begin": (1.e-2, "density below which the sponge begins"),
        "sponge.sponge_rho_full": (1.e-3, "density below which the sponge is fully on"),
        "sponge.sponge_timescale": (1.e-2, "time scale of the damping"),
        "particles.do_particles": (0, ""),
    },
    "advection": {
        "driver.cfl": (0.8, "advective CFL number"),
        "advection.u": (1.0, "advective velocity in x"),
        "advection.v": (1.0, "advective velocity in y"),
        "advection.limiter": (2, "0 none, 1 second-order MC, 2 fourth-order MC"),
        "particles.do_particles": (0, "not supported"),
        "particles.particle_generator": ("grid", ""),
    },
    "diffusion": {
        "driver.cfl": (0.8, "diffusion CFL number (may exceed 1: the update is implicit)"),
        "diffusion.k": (1.0, "conductivity"),
        "diffusion.mg_split_n": (1024, "multi-GPU runs (extension): multigrid levels with at least this many columns "
                                       "are split into x-slabs, coarser ones are replicated"),
    },
    "lm_atm": {
        "driver.cfl": (0.8, ""),
        "lm-atmosphere.limiter": (2, "0 none, 1 second-order MC, 2 fourth-order MC"),
        "lm-atmosphere.proj_type": (2, "what is projected: 1 includes the -Gp term in U*"),
        "lm-atmosphere.grav": (-2.0, "gravitational acceleration along y"),
        "eos.gamma": (1.4, "pres = rho ener (gamma - 1)"),
    },
    "burgers": {
        "driver.cfl": (0.8, "advective CFL number"),
        "advection.limiter": (2, "0 none, 1 second-order MC, 2 fourth-order MC"),
        "particles.do_particles": (0, "not supported"),
        "particles.particle_generator": ("grid", ""),
    },
    "incompressible": {
        "driver.cfl": (0.8, ""),
        "incompressible.limiter": (2, "0 none, 1 second-order MC, 2 fourth-order MC"),
        "incompressible.proj_type": (2, "what is projected: 1 includes the -Gp term in U*"),
        "incompressible.mg_split_n": (1024, "multi-GPU runs (extension): multigrid levels with at least this many columns "
                                            "are split into x-slabs, coarser ones are replicated"),
        "particles.do_particles": (0, "not supported"),
        "particles.particle_generator": ("grid", ""),
    },
}
