"""The reference's own end-to-end regression tests (tests_cpp/*_explicit, ctest + compare.py) as data: problem, mesh,
stepper, step size, step count, what compare.py checks and with which tolerance.  The gold files themselves are
converted to tests/golden/refgold/refgold.npz by tests/golden/make_refgold.py (the reference is not on the GPU box).

stencil per scheme as in the reference's CMake files: firstorder -s 3, weno3 -s 5, weno5 -s 7.
check kinds: "state" (whole final state), "state@100+150" (states after 100 and 150 steps, concatenated), "rho" / "h" (dof 0 of every cell), "p" (pressure from the conserved state,
gamma = 1.4), "rho_linf" (L-inf error of the density against the analytic solution == a constant asserted by
compare.py with math.isclose, rel_tol 1e-9).
"""
SCHEMES = {"firstorder": ("FirstOrder", 3), "weno3": ("Weno3", 5), "weno5": ("Weno5", 7)}


def _c(ref_dir, schemes, enum, n, bounds, periodic, stepper, dt, nsteps, checks, rtol, atol, ic=1, factory=None,
       subdirs=None):
    return dict(ref_dir=ref_dir, schemes=schemes, enum=enum, n=n, bounds=bounds, periodic=periodic, stepper=stepper,
                dt=dt, nsteps=nsteps, checks=checks, rtol=rtol, atol=atol, ic=ic, factory=factory, subdirs=subdirs)


ALL3 = ("firstorder", "weno3", "weno5")
CASES = {
    # tests_cpp/eigen_1d_euler_sod_explicit/{main.cc,compare.py,test.cmake}
    "sod1d": _c("eigen_1d_euler_sod_explicit", ALL3, ("Euler1d", "Sod"), [100], [-0.5, 0.5], (), "rk4", 0.001, 100,
                {"state": "gold.txt"}, 1e-9, 1e-11),
    "lax1d": _c("eigen_1d_euler_lax_explicit", ALL3, ("Euler1d", "Lax"), [100], [-5.0, 5.0], (), "ssprk3", 0.0001, 100,
                {"state": "gold.txt"}, 1e-9, 1e-11),
    # compare.py reshapes the 4 observer snapshots (steps 0,50,100,150; 750 dofs each) with nx=500, so what it compares
    # with the gold is [state@100, state@150] of the 250-cell run (verified against the compiled reference: exact)
    "shuosher1d": _c("eigen_1d_euler_shu_osher_explicit", ALL3, ("Euler1d", "ShuOsher"), [250], [-5.0, 5.0], (), "ssprk3",
                     0.001, 150, {"state@100+150": "gold.txt"}, 1e-9, 1e-11),
    "advection1d": _c("eigen_1d_linear_advection_default_velocity_explicit", ALL3, ("Advection1d", "PeriodicLinear"), [200],
                      [-1.0, 1.0], ("x",), "rk4", 0.001, 100, {"state": "gold.txt"}, 1e-9, 1e-11),
    "riemann2d": _c("eigen_2d_euler_riemann_explicit", ALL3, ("Euler2d", "Riemann"), [20, 20], [0.0, 1.0, 0.0, 1.0], (),
                    "ssprk3", 0.01, 60, {"rho": "rho_gold.txt", "p": "p_gold.txt"}, 1e-10, 1e-12, ic=2),
    "sedov2d": _c("eigen_2d_euler_sedov_explicit", ALL3, ("Euler2d", "SedovFull"), [25, 25], [-0.5, 0.5, -0.5, 0.5], (),
                  "ssprk3", 0.001, 50, {"p": "p_gold.txt"}, 1e-10, 1e-12),
    "sedovsym2d": _c("eigen_2d_euler_sedov_symmetry_explicit", ALL3, ("Euler2d", "SedovSymmetry"), [25, 25],
                     [0.0, 0.5, 0.0, 0.5], (), "ssprk3", 0.0005, 100, {"state": "final_state_gold.txt"}, 1e-10, 1e-12),
    "normalshock2d": _c("eigen_2d_euler_normal_shock_explicit", ALL3, ("Euler2d", "NormalShock"), [40, 20],
                        [0.0, 2.0, 0.0, 1.0], (), "ssprk3", 0.001, 100, {"rho": "rho_gold.txt"}, 1e-10, 1e-12),
    "crossshock2d": _c("eigen_2d_euler_cross_shock_explicit", ALL3, ("Euler2d", "CrossShock"), [40, 20],
                       [0.0, 2.0, 0.0, 1.0], (), "ssprk3", 0.001, 100, {"rho": "rho_gold.txt"}, 1e-10, 1e-12,
                       factory="cross_shock"),
    "dmr2d": _c("eigen_2d_euler_double_mach_reflection_explicit", ("firstorder", "weno3"), ("Euler2d", "DoubleMachReflection"),
                [60, 15], [0.0, 4.0, 0.0, 1.0], (), "ssprk3", 0.001, 150, {"rho": "rho_gold.txt"}, 1e-10, 1e-12),
    "kh2d": _c("eigen_2d_euler_kelvin_helmholtz_explicit_short", ALL3, ("Euler2d", "KelvinHelmholtz"), [25, 25],
               [-5.0, 5.0, -5.0, 5.0], ("x", "y"), "rk4", 0.010439892262204077, 100, {"p": "p_gold.txt"}, 1e-10, 1e-12),
    "smooth2d": _c("eigen_2d_euler_smooth_explicit", ALL3, ("Euler2d", "PeriodicSmooth"), [25, 25], [-1.0, 1.0, -1.0, 1.0],
                   ("x", "y"), "rk4", 0.01, 200,
                   {"rho_linf": {"firstorder": 0.19629186956424138, "weno3": 0.09177922458156529,
                                 "weno5": 0.000723019713426809}}, 1e-9, 0.0),
    "swe2d": _c("eigen_2d_swe_slip_wall_explicit", ALL3, ("Swe2d", "SlipWall"), [25, 25], [-5.0, 5.0, -5.0, 5.0], (),
                "rk4", 0.01, 200, {"h": "h_gold.txt"}, 1e-10, 1e-12),
    "swe2d_ic2": _c("eigen_2d_swe_slip_wall_explicit_ic2", ALL3, ("Swe2d", "SlipWall"), [25, 25], [-5.0, 5.0, -5.0, 5.0], (),
                    "rk4", 0.01, 400, {"h": "h_gold.txt"}, 1e-10, 1e-12, ic=2),
    "burgers2d": _c("eigen_2d_burgers_periodic_explicit", ALL3, ("AdvectionDiffusion2d", "BurgersPeriodic"), [20, 20],
                    [-1.0, 1.0, -1.0, 1.0], ("x", "y"), "rk4", 0.01, 200, {"state": "gold.txt"}, 1e-10, 1e-12),
    "advdiffreac2d": _c("eigen_2d_advdiffreac_probA_explicit", ALL3, ("AdvectionDiffusionReaction2d", "ProblemA"), [18, 18],
                        [0.0, 1.0, 0.0, 1.0], (), "rk4", 0.01, 500, {"h": "gold.txt"}, 1e-10, 1e-12),
    "grayscott2d": _c("eigen_2d_gray_scott_explicit", ("firstorder",), ("DiffusionReaction2d", "GrayScott"), [20, 20],
                      [-1.25, 1.25, -1.25, 1.25], ("x", "y"), "rk4", 0.8, 2000, {"state": "gold.txt"}, 1e-5, 1e-8,
                      subdirs=False),
    "diffreac2d": _c("eigen_2d_diffusion_reaction_explicit", ("firstorder",), ("DiffusionReaction2d", "ProblemA"), [35, 35],
                     [0.0, 1.0, 0.0, 1.0], (), "ssprk3", 0.001, 100, {"state": "gold.txt"}, 1e-10, 1e-12, subdirs=False),
    "diffreac1d": _c("eigen_1d_diffusion_reaction_explicit", ("firstorder",), ("DiffusionReaction1d", "ProblemA"), [100],
                     [0.0, 1.0], (), "rk4", 0.001, 1000, {"state": "gold.txt"}, 1e-9, 1e-11, factory="diffreac1d_A",
                     subdirs=False),
    "sedovsym3d_equal": _c("eigen_3d_euler_sedov_symmetry", ("firstorder", "weno3"), ("Euler3d", "SedovSymmetry"),
                           [20, 20, 20], [0.0, 0.4, 0.0, 0.4, 0.0, 0.4], (), "ssprk3", 0.0001, 250,
                           {"state": "gold_state.txt"}, 1e-8, 1e-10, subdirs="{scheme}_nxnynz_equal"),
    "sedovsym3d_notequal": _c("eigen_3d_euler_sedov_symmetry", ("firstorder", "weno3"), ("Euler3d", "SedovSymmetry"),
                              [20, 25, 28], [0.0, 0.4, 0.0, 0.4, 0.0, 0.4], (), "ssprk3", 0.0001, 250,
                              {"state": "gold_state.txt"}, 1e-8, 1e-10, subdirs="{scheme}_nxnynz_not_equal"),
    "smooth3d": _c("eigen_3d_euler_smooth_short", ("firstorder", "weno3"), ("Euler3d", "PeriodicSmooth"), [8, 8, 8],
                   [-1.0, 1.0, -1.0, 1.0, -1.0, 1.0], ("x", "y", "z"), "rk4", 0.01, 200,
                   {"rho_linf": {"firstorder": 0.18477590666454702, "weno3": 0.1844736236322082}}, 1e-9, 0.0),
}


def gold_key(case, scheme, check):
    return "%s/%s/%s" % (case, scheme, check)
