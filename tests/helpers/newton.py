"""Tiny load-stepping Newton driver (dense solve) used to reproduce the reference's integration-test
known answers with any evaluate+assemble backend (oracle on CPU, CUDA path on GPU).

Mirrors what the reference's time integrator does around the hot path for DYNAMICTYPE Statics
(src/structure_new/src/implicit/4C_structure_new_impl_statics.cpp:93-98): residual
f_int(u) - f_ext(t), Jacobian K_T, Dirichlet rows eliminated, converge on TOLRES and TOLDISP.
"""
import numpy as np

from .meshes import csr_to_dense
from .neumann import neumann_load


def solve_case(case, evaluate, max_iter=30, update_prestress=None, recover=None, external=None, update_step=None):
    """evaluate(u_flat) -> (rowptr, colind, values, fint). Returns converged displacement (nn,3).

    recover(du_flat): element action struct_calc_recover after every Newton update with the Newton
    direction as "residual displacement" (Solid::ModelEvaluator::Structure::run_post_compute_x / run_recover,
    model_evaluator_structure.cpp:1482-1509): condensed element variables (EAS alpha) follow the step.

    external(u_flat, t, values) -> (fext, values'): deformation dependent loads (orthopressure with LOADLIN:
    apply_force_stiff_external adds the load stiffness to the stiffness matrix of the iteration,
    model_evaluator_structure.cpp:474-516); default is the dead load of the case.

    update_step(u_flat): end of a converged time step (update_step_element -> struct_calc_update_istep:
    history materials commit their current state, Mat::...::update()).

    update_prestress(u_flat): PRESTRESS MULF.  While the time reached by a step is <= PRESTRESSTIME
    (the state has already been advanced when update_step_element runs; confirmed by the known answers:
    PRESTRESSTIME 1 with TIMESTEP 0.5 gives two MULF steps) the converged step ends with the element
    action struct_update_prestress and the displacements are reset to zero (src/structure_new/src/model_evaluator/4C_structure_new_model_evaluator_structure.cpp:
    1608-1642, src/structure_new/src/implicit/4C_structure_new_impl_prestress.cpp:94-108)."""
    nn = len(case["coords"])
    fixed = np.zeros(3 * nn, bool)
    for n in case["dirichlet_nodes"]:
        fixed[3 * n:3 * n + 3] = True
    fixed[case.get("dirichlet_dofs", [])] = True
    combine_or = case.get("normcombi", "And") == "Or"  # NORMCOMBI_RESFDISP
    free = ~fixed
    u = np.zeros(3 * nn)
    t = 0.0
    for _ in range(case["numstep"]):
        t += case["timestep"]
        fext = neumann_load(case, t) if external is None else None
        for it in range(max_iter):
            rowptr, colind, values, fint = evaluate(u)
            if external is not None:
                fext, values = external(u, t, values)
            res = fint - fext
            du = np.zeros(3 * nn)
            if 3 * nn > 1500:
                import scipy.sparse as sp
                from scipy.sparse.linalg import spsolve
                K = sp.csr_matrix((values, colind, rowptr), shape=(3 * nn, 3 * nn))
                du[free] = spsolve(K[free][:, free].tocsc(), -res[free])
            else:
                K = csr_to_dense(rowptr, colind, values, 3 * nn)
                du[free] = np.linalg.solve(K[np.ix_(free, free)], -res[free])
            u += du
            if recover is not None:
                recover(du)
            rowptr, colind, values, fint = evaluate(u)
            if external is not None:
                fext, _ = external(u, t, None)
            rnorm = np.linalg.norm((fint - fext)[free])
            ok_r, ok_d = rnorm < case["tolres"], np.linalg.norm(du) < case["toldisp"]
            if (ok_r or ok_d) if combine_or else (ok_r and ok_d):
                break
        else:
            raise RuntimeError(f"Newton did not converge at t={t}: |r|={rnorm}")
        if update_step is not None:
            update_step(u)
        if update_prestress is not None and t <= case["prestresstime"] + 1.0e-15:
            update_prestress(u)
            u = np.zeros(3 * nn)
    return u.reshape(nn, 3)


def solve_prescribed_displacement_case(case, evaluate, max_iter=30, recover=None, update_step=None, trace=None,
                                       predictor=None):
    """Statics driven by DESIGN ... DIRICH CONDITIONS with time functions only (no external load): every
    step sets the prescribed dofs to VAL * FUNCT(t) (Core::FE::evaluate_dirichlet) and iterates the free
    ones on f_int = 0.  case["dirichlet"]: [{nodes, ONOFF, VAL, FUNCT[, level]}], case["functs"][id] =
    linearinterpolation {TIMES, VALUES}.  level (0 surface, 1 line, 2 point): a lower-dimensional entity overrides the
    higher ones for ALL dofs of its nodes, entities of one level add up (Core::FE::Utils::Dbc::read_dirichlet_condition,
    src/core/fem/src/condition/4C_fem_condition_utils... hierarchical order).
    recover(du_flat) / update_step(u_flat): as in solve_case (EAS recover after every Newton update with the iteration
    increment; history materials commit at the end of a converged step).
    predictor "constvel": PREDICT ConstVel -- the free dofs start a step from u_n + (u_n - u_(n-1)), the statics form of
    the constant-velocity predictor (src/structure_new/src/predict/4C_structure_new_predict_constdisvelaccpress.cpp)."""
    nn = len(case["coords"])
    fixed = np.zeros(3 * nn, bool)
    governs = -np.ones((nn, 3), int)  # index of the condition that prescribes the dof
    level = -np.ones(nn, int)
    for ic, c in enumerate(case["dirichlet"]):
        lv = c.get("level", 0)
        for n in c["nodes"]:
            if lv > level[n]:
                governs[n, :] = -1
                level[n] = lv
            if lv == level[n]:
                for d in range(3):
                    if c["ONOFF"][d]:
                        governs[n, d] = ic
    fixed = (governs >= 0).ravel()
    free = ~fixed
    u = np.zeros(3 * nn)
    u_old = u.copy()
    t = 0.0
    for step in range(case["numstep"]):
        t += case["timestep"]
        if predictor == "constvel":
            u, u_old = u + (u - u_old), u.copy()
        du_dbc = np.zeros(3 * nn)
        for n in range(nn):
            for d in range(3):
                ic = governs[n, d]
                if ic < 0:
                    continue
                c = case["dirichlet"][ic]
                f = case["functs"][str(c["FUNCT"][d])] if c["FUNCT"][d] else None
                fac = 1.0 if f is None else float(np.interp(t, f["TIMES"], f["VALUES"]))
                du_dbc[3 * n + d] = c["VAL"][d] * fac - u[3 * n + d]
        u = u + du_dbc
        for it in range(max_iter):
            rowptr, colind, values, fint = evaluate(u)
            du = np.zeros(3 * nn)
            if 3 * nn > 1500:
                import scipy.sparse as sp
                from scipy.sparse.linalg import spsolve
                K = sp.csr_matrix((values, colind, rowptr), shape=(3 * nn, 3 * nn))
                du[free] = spsolve(K[free][:, free].tocsc(), -fint[free])
            else:
                K = csr_to_dense(rowptr, colind, values, 3 * nn)
                du[free] = np.linalg.solve(K[np.ix_(free, free)], -fint[free])
            u += du
            if recover is not None:
                recover(du)
            rowptr, colind, values, fint = evaluate(u)
            rn = np.linalg.norm(fint[free])
            if trace is not None:
                trace.append((step, it, rn, np.linalg.norm(du)))
            if rn < case["tolres"] and np.linalg.norm(du) < case["toldisp"]:
                break
        else:
            raise RuntimeError(f"Newton did not converge at t={t}: |r|={np.linalg.norm(fint[free])}")
        if update_step is not None:
            update_step(u)
    return u.reshape(nn, 3)


def solve_genalpha_case(case, evaluate, mass_values, recover=None, update_step=None, max_iter=25, trace=None):
    """Generalised-alpha time integration around the hot path, as structure_new runs it for DYNAMICTYPE GenAlpha
    (the default of STRUCTURAL DYNAMIC) with GENAVG TrLike, no damping and the ConstDis predictor:
      coefficients from RHO_INF        4C_structure_new_impl_genalpha.cpp:59,175-190
      Newmark update of v, a           :205-250, 460-510
      residual  (1-a_f)(f_int - f_ext)_{n+1} + a_f (f_int - f_ext)_n + M ((1-a_m) a_{n+1} + a_m a_n)
                                        :264-262,326-344; model_evaluator_structure.cpp:277-300,1539-1570
      Jacobian  (1-a_f) K + (1-a_m)/(beta dt^2) M                                            :346-360
      consistent initial acceleration  M a_0 = f_ext(0) - f_int(0)  (equilibrate_initial_state)
    evaluate(u_flat) -> (rowptr, colind, K values, f_int); mass_values: CSR values of M on the same graph.
    recover(du) / update_step(u): as in solve_case.  Returns the displacement (nn,3) after numstep steps."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import splu
    nn = len(case["coords"])
    ndof = 3 * nn
    fixed = np.zeros(ndof, bool)
    for n in case.get("dirichlet_nodes", []):
        fixed[3 * n:3 * n + 3] = True
    fixed[case.get("dirichlet_dofs", [])] = True
    free = np.nonzero(~fixed)[0]
    rho = case["genalpha"]["rho_inf"]
    am, af = (2.0 * rho - 1.0) / (rho + 1.0), rho / (rho + 1.0)
    beta, gamma = 0.25 * (1.0 - am + af) ** 2, 0.5 - am + af
    dt = case["timestep"]
    d, v = np.zeros(ndof), np.zeros(ndof)
    rowptr, colind, Kv, fint_n = evaluate(d)
    M = sp.csr_matrix((mass_values, colind, rowptr), shape=(ndof, ndof))
    Mff = M[free][:, free].tocsc()
    t = 0.0
    fext_n = neumann_load(case, t)
    a = np.zeros(ndof)
    a[free] = splu(Mff).solve((fext_n - fint_n)[free])
    for step in range(case["numstep"]):
        t += dt
        fext = neumann_load(case, t)
        const_vel = (beta - gamma) / beta * v + (2.0 * beta - gamma) * dt / (2.0 * beta) * a - gamma / (beta * dt) * d
        const_acc = (2.0 * beta - 1.0) / (2.0 * beta) * a - v / (beta * dt) - d / (beta * dt * dt)
        dn = d.copy()        # predictor ConstDis
        for it in range(max_iter):
            rowptr, colind, Kv, fint = evaluate(dn)
            an = const_acc + dn / (beta * dt * dt)
            an[fixed] = 0.0
            R = (1.0 - af) * (fint - fext) + af * (fint_n - fext_n) + M @ ((1.0 - am) * an + am * a)
            rnorm = np.linalg.norm(R[free])
            if it > 0 and np.linalg.norm(du) < 1e-13 * max(1.0, np.linalg.norm(dn)) + 1e-16:
                break
            K = sp.csr_matrix((Kv, colind, rowptr), shape=(ndof, ndof))
            Keff = ((1.0 - af) * K + (1.0 - am) / (beta * dt * dt) * M)[free][:, free].tocsc()
            du = np.zeros(ndof)
            du[free] = splu(Keff).solve(-R[free])
            dn += du
            if recover is not None:
                recover(du)
        else:
            raise RuntimeError(f"GenAlpha Newton did not converge at t={t}: |r|={rnorm}")
        vn = const_vel + gamma / (beta * dt) * dn
        an = const_acc + dn / (beta * dt * dt)
        vn[fixed], an[fixed] = 0.0, 0.0
        d, v, a, fint_n, fext_n = dn, vn, an, fint, fext
        if update_step is not None:
            update_step(d)
        if trace is not None:
            trace.append((t, d.copy()))
    return d.reshape(nn, 3)
