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


def solve_prescribed_displacement_case(case, evaluate, max_iter=30):
    """Statics driven by DESIGN SURF DIRICH CONDITIONS with time functions only (no external load): every
    step sets the prescribed dofs to VAL * FUNCT(t) (Core::FE::evaluate_dirichlet) and iterates the free
    ones on f_int = 0.  case["dirichlet"]: [{nodes, ONOFF, VAL, FUNCT}], case["functs"][id] =
    linearinterpolation {TIMES, VALUES}."""
    nn = len(case["coords"])
    fixed = np.zeros(3 * nn, bool)
    for c in case["dirichlet"]:
        for n in c["nodes"]:
            for d in range(3):
                if c["ONOFF"][d]:
                    fixed[3 * n + d] = True
    free = ~fixed
    u = np.zeros(3 * nn)
    t = 0.0
    for _ in range(case["numstep"]):
        t += case["timestep"]
        for c in case["dirichlet"]:
            for d in range(3):
                if not c["ONOFF"][d]:
                    continue
                f = case["functs"][str(c["FUNCT"][d])] if c["FUNCT"][d] else None
                fac = 1.0 if f is None else float(np.interp(t, f["TIMES"], f["VALUES"]))
                for n in c["nodes"]:
                    u[3 * n + d] = c["VAL"][d] * fac
        for it in range(max_iter):
            rowptr, colind, values, fint = evaluate(u)
            K = csr_to_dense(rowptr, colind, values, 3 * nn)
            du = np.zeros(3 * nn)
            du[free] = np.linalg.solve(K[np.ix_(free, free)], -fint[free])
            u += du
            rowptr, colind, values, fint = evaluate(u)
            if np.linalg.norm(fint[free]) < case["tolres"] and np.linalg.norm(du) < case["toldisp"]:
                break
        else:
            raise RuntimeError(f"Newton did not converge at t={t}: |r|={np.linalg.norm(fint[free])}")
    return u.reshape(nn, 3)
