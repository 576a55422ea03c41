"""numpy mirror of k_svsh_solve (intrinsic3d_b200/csrc/i3d_lighting.cuh): the Ceres trust-region loop of
LightingSVSH::estimate on the REDUCED system (per-subvolume 9x9 normal equations + ring Laplacian).

Test helper only.  tests/test_lighting.py checks it against the oracle's explicit-row restatement, which pins the
algebra the CUDA kernel relies on (cost / gradient / model change as functions of (H, g, c), the factor 2 of the
directed pairs, Jacobi scaling and the LM diagonal from the block diagonals)."""
import numpy as np


def reduce_rows(S, rows):
    """SHDataCost rows -> what k_svsh_accumulate sums per subvolume: H [S,9,9], g [S,9], c, sum_w, count."""
    sub, j, lum, w = rows["sub"], rows["j"], rows["lum"], rows["w"]
    H = np.zeros((S, 9, 9))
    g = np.zeros((S, 9))
    np.add.at(H, sub, w[:, None, None] * j[:, :, None] * j[:, None, :])
    np.add.at(g, sub, (w * lum)[:, None] * j)
    return H, g, float(np.sum(w * lum * lum)), float(np.sum(w)), int(len(w))


def solve_reduced(H_raw, g_raw, c_raw, sum_w, nbr, P):
    """nbr: [S,6] neighbour subvolume ids (-1 = none), ring order.  Returns (x [S,9], info dict)."""
    S = H_raw.shape[0]
    M = 9 * S
    deg = (nbr >= 0).sum(1).astype(np.float64)
    n_pairs = float(deg.sum())
    data_loss = 1.0 / sum_w if sum_w > 0 else 1.0
    H = H_raw * data_loss
    g = (g_raw * data_loss).reshape(M)
    c0 = c_raw * data_loss
    wr2 = 2.0 * (P.lambda_reg / n_pairs if n_pairs > 0 else 0.0)

    def apply_A(v):
        v = v.reshape(S, 9)
        out = np.einsum("sij,sj->si", H, v)
        nb = np.zeros_like(v)
        for d in range(6):
            m = nbr[:, d] >= 0
            nb[m] += v[nbr[m, d]]
        return (out + wr2 * (deg[:, None] * v - nb)).reshape(M)

    colsq = (np.einsum("sii->si", H) + wr2 * deg[:, None]).reshape(M)
    scale = 1.0 / (1.0 + np.sqrt(colsq))
    diag = np.clip(colsq * scale * scale, P.min_lm_diagonal, P.max_lm_diagonal)
    x = np.zeros(M)
    gU = -g.copy()
    gmax = np.abs(g).max() if M else 0.0
    cost = 0.5 * c0
    info = dict(cost_initial=cost, lm_iterations=0, num_successful_steps=0, cg_iterations_total=0, termination=1)
    x_norm = 0.0
    radius, decrease = P.initial_trust_region_radius, 2.0
    invalid = 0
    it = 0
    while True:
        if it >= P.max_iterations:
            info["termination"] = 1
            break
        if gmax <= P.gradient_tolerance or radius <= P.min_trust_region_radius:
            info["termination"] = 0
            break
        it += 1
        D2 = diag / radius
        b = scale * gU
        sc = scale.reshape(S, 9)
        B = H + np.eye(9)[None] * (wr2 * deg)[:, None, None]
        B = B * sc[:, :, None] * sc[:, None, :] + np.eye(9)[None] * D2.reshape(S, 9)[:, :, None]
        Minv = np.linalg.inv(B)
        xs = np.zeros(M)
        r = b.copy()
        cg_it = 0
        failed = False
        if np.sqrt(b @ b) != 0.0:
            rho, Q0 = 1.0, 0.0
            p = None
            cg_it = 1
            while True:
                z = np.einsum("sij,sj->si", Minv, r.reshape(S, 9)).reshape(M)
                last_rho, rho = rho, r @ z
                if rho == 0.0 or np.isinf(rho):
                    failed = True
                    break
                if cg_it == 1:
                    p = z
                else:
                    beta = rho / last_rho
                    if beta == 0.0 or np.isinf(beta):
                        failed = True
                        break
                    p = z + beta * p
                q = scale * apply_A(scale * p) + D2 * p
                pq = p @ q
                if pq <= 0.0 or np.isinf(pq):
                    break
                alpha = rho / pq
                xs = xs + alpha * p
                if cg_it % P.residual_reset_period == 0:
                    r = b - (scale * apply_A(scale * xs) + D2 * xs)
                else:
                    r = r - alpha * q
                Q1 = -(xs @ (b + r))
                zeta = cg_it * (Q1 - Q0) / Q1
                if zeta < P.eta and cg_it >= P.min_linear_solver_iterations:
                    break
                Q0 = Q1
                if cg_it >= P.max_linear_solver_iterations:
                    break
                cg_it += 1
        info["cg_iterations_total"] += cg_it
        s = -xs
        delta = scale * s
        valid = (not failed) and bool(np.all(np.isfinite(s)))
        model = 0.0
        if valid:
            Ad = apply_A(delta)
            model = -(s @ b + 0.5 * (delta @ Ad))
            valid = model > 0.0
        if not valid:
            invalid += 1
            if invalid >= P.max_consecutive_invalid_steps:
                info["termination"] = 2
                break
            radius *= 0.5
            continue
        invalid = 0
        cx = x + delta
        cand = 0.5 * (cx @ (gU + Ad - g) + c0)
        step_norm = np.sqrt(((x - cx) ** 2).sum())
        if step_norm <= P.parameter_tolerance * (x_norm + P.parameter_tolerance):
            info["termination"] = 0
            break
        change = cost - cand
        if abs(change) <= P.function_tolerance * cost:
            info["termination"] = 0
            break
        rho_q = change / model
        if rho_q > P.min_relative_decrease:
            x, gU, cost = cx, gU + Ad, cand
            x_norm = np.sqrt(x @ x)
            gmax = np.abs(gU).max()
            u = 2.0 * rho_q - 1.0
            radius = min(P.max_trust_region_radius, radius / max(1.0 / 3.0, 1.0 - u * u * u))
            decrease = 2.0
            info["num_successful_steps"] += 1
        else:
            radius /= decrease
            decrease *= 2.0
    info.update(lm_iterations=it, cost_final=cost, trust_region_radius=radius)
    return x.reshape(S, 9), info
