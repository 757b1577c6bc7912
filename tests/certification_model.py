"""numpy model of the certified float pass (botlab_b200/csrc/mcl_device.cuh: score_beam_fast) -- TEST INFRASTRUCTURE.

It restates, in float32 numpy, the arithmetic and the certification rules of the first sensor pass, and the error budget
of mcl_engine.cu: fast_plan.  tests/test_certification_model.py uses it on the CPU to check the DESIGN (DESIGN.md
section 5) independently of the CUDA code: whatever the model calls "certain" must equal the oracle's per-ray score,
with an adversarial perturbation of the sine/cosine up to the SFU error bound the budget assumes.  It is also the place
to try changes to the certification (tighter budgets, other tests) before spending GPU time.  Keep it in step with
score_beam_fast / fast_plan when those change."""
import numpy as np

F = np.float32
U = 5.9604644775390625e-08          # float unit roundoff 2^-24
TRIG_ERR = 2.0e-6                   # kFastTrigErr


def fma32(a, b, c):
    """float32 fused multiply-add: the product of two floats is exact in double; one rounding to double, one to float
    (double rounding can differ from a true FMA in ~2^-29 of the cases, far below what the budget resolves)."""
    return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(F) if np.ndim(b) == 0 else \
        (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


class Plan:
    """fast_plan + plan_set_window for a window [x0, x0+w) x [y0, y0+h) of global cells.  The float model works in
    WINDOW-RELATIVE cell coordinates (its three coordinate roundings then happen at the window's magnitude, not the
    map's) with as many fixed-point fractional bits as the window's extent leaves room for."""

    def __init__(self, grid, ranges, thetas, ratios, min_range, x0, y0, w, h):
        cpm = float(grid.cells_per_meter)
        valid = ranges > F(min_range)
        finite = np.isfinite(ranges[valid])
        rc_max = float(ranges[valid][finite].max()) * cpm if finite.any() else 0.0
        rho_lo, rho_hi = float(ratios[valid].min()), float(ratios[valid].max())
        rho_max = max(abs(rho_lo), abs(rho_hi))
        cm = max(x0 + w, y0 + h) + 1.0
        we = max(w, h) + rc_max + 8.0            # magnitude bound of window-relative coordinates the budget covers
        self.enabled = (valid.any() and min_range * cpm >= 2.5 and float(np.abs(thetas[valid]).max()) <= 6.3
                        and rho_lo >= -1.0 and rho_hi <= 2.0 and w >= 3 and h >= 3 and cm <= 4090 and x0 >= -4 and y0 >= -4
                        and max(w, h) + 8 < 4096)
        if not self.enabled:
            return
        wd = max(w, h) + 8                       # only coordinates inside the window have their fixed-point bits used
        self.fb = fb = 12 if wd < 1024 else (11 if wd < 2048 else 10)
        xm = cm / cpm + max(abs(grid.origin_x), abs(grid.origin_y))
        shift = 64.0
        ce = cm + rc_max
        e_ref = cpm * U * xm + 2 * U * ce + 2 * U * rc_max + rc_max * (20 * U + 1.2e-7) + 1e-9
        e_apx = 3 * U * we + (1 + 2 * rho_max) * U * shift + 2 * U * rc_max + \
            rc_max * ((np.pi * (3 * rho_max + 1) + 9.5) * U + TRIG_ERR)
        self.eps = 1.25 * (e_ref + e_apx) + 1e-6
        one = 1 << fb
        k = int(np.ceil(one * self.eps + 0.5))
        if k > one // 32:
            self.enabled = False
            return
        kb = 1
        while kb < k:
            kb *= 2
        self.kb = kb
        self.fmask = (one - 1) & ~(2 * kb - 1)
        self.magic_base = F(1.5 * 2.0 ** (23 - fb))
        self.magic = self.magic_base + F(kb) / F(one)
        self.mbk = int(self.magic_base.view(np.int32)) >> fb
        self.t_dir = F(3.0 * (1.0 + self.eps) + 4.0 * U * rc_max + 1e-4)
        self.t_dir_neg = F(5.0 * (1.0 + self.eps) + 4.0 * U * rc_max + 1e-4)
        self.rho_lo, self.rho_hi, self.max_shift, self.coord_hi = F(rho_lo), F(rho_hi), F(shift), F(cm - 1.0)
        # window-dependent part (plan_set_window): everything below is in window-relative cells
        self.x0, self.y0 = x0, y0
        x2_min = 3.0 * self.eps + 1e-3
        self.x2_lo_x, self.x2_lo_y = F(x2_min - x0), F(x2_min - y0)
        self.gmid_x, self.ghalf_x = F(0.5 * (grid.width - 1) - x0), F(0.5 * (grid.width + 3) + self.eps + 1e-3)
        self.gmid_y, self.ghalf_y = F(0.5 * (grid.height - 1) - y0), F(0.5 * (grid.height + 3) + self.eps + 1e-3)
        lcx, hcx = max(1, -x0), w - 1
        lcy, hcy = max(1, -y0), h - 1
        slack = (self.magic - self.magic_base) + F(0.5) / F(one)
        self.mid_x, self.half_x = F(0.5) * F(lcx + hcx), F(0.5) * F(hcx - lcx) - slack
        self.mid_y, self.half_y = F(0.5) * F(lcy + hcy), F(0.5) * F(hcy - lcy) - slack


def cloud_window(grid, cloud, ranges, min_range):
    """The single-tile window of mcl_engine.cu: run_score: bounding box of poses and parents +- (max range in cells + 3),
    clipped to the grid plus a 2-cell margin, x0 aligned down to a multiple of 4."""
    cpm = float(grid.cells_per_meter)
    valid = ranges > F(min_range)
    reach = float(ranges[valid][np.isfinite(ranges[valid])].max()) * cpm + 3.0
    xs = np.concatenate([cloud["pose"]["x"], cloud["parent_pose"]["x"]]).astype(np.float64)
    ys = np.concatenate([cloud["pose"]["y"], cloud["parent_pose"]["y"]]).astype(np.float64)
    cx0 = np.floor((xs.min() - grid.origin_x) * cpm - reach); cx1 = np.ceil((xs.max() - grid.origin_x) * cpm + reach)
    cy0 = np.floor((ys.min() - grid.origin_y) * cpm - reach); cy1 = np.ceil((ys.max() - grid.origin_y) * cpm + reach)
    x0, y0 = int(max(cx0, -2)), int(max(cy0, -2))
    x1, y1 = int(min(cx1, grid.width + 1)), int(min(cy1, grid.height + 1))
    x0 = (x0 & ~3) if x0 >= 0 else -(((-x0) + 3) & ~3)
    return x0, y0, x1 - x0 + 1, y1 - y0 + 1


def derive_fast_map(cells):
    """derive_fast_map_kernel: positive cells unchanged; non-positive ones -1 if their 5x5 neighbourhood holds nothing
    positive, else 0."""
    pos = np.pad(cells > 0, 2)
    h, w = cells.shape
    near = np.zeros((h, w), bool)
    for dy in range(5):
        for dx in range(5):
            near |= pos[dy:dy + h, dx:dx + w]
    out = np.where(cells > 0, cells, np.where(near, 0, -1)).astype(np.int8)
    return out


def fast_pass(grid, plan, particle, ranges, thetas, ratios, min_range, fast_cells, rng, interp=True):
    """One particle, all valid beams.  Returns (half_unit_scores, certain) per valid beam, in scan order.
    rng perturbs the sine/cosine by up to +-1.3e-6 (the measured SFU error), adversarially for the certification."""
    valid = ranges > F(min_range)
    r, th, rho = ranges[valid], thetas[valid], ratios[valid].astype(F)
    gx, gy, cpm_d = np.float64(F(grid.origin_x)), np.float64(F(grid.origin_y)), np.float64(F(grid.cells_per_meter))
    xa, ya, tha = (F(particle["pose"][k]) for k in ("x", "y", "theta"))
    xb, yb, thb = (F(particle["parent_pose"][k]) for k in ("x", "y", "theta"))
    sh_x, sh_y = np.float64(plan.x0), np.float64(plan.y0)
    if interp:
        gsx, gsy = (np.float64(xb) - gx) * cpm_d, (np.float64(yb) - gy) * cpm_d
        dsx, dsy = F(np.float64(F(xa - xb)) * cpm_d), F(np.float64(F(ya - yb)) * cpm_d)
        d = np.float64(tha) - np.float64(thb)
        if abs(d) > np.pi:
            d += -2 * np.pi if d > 0 else 2 * np.pi
        th0, dth = thb, F(d)
    else:
        gsx, gsy = (np.float64(xa) - gx) * cpm_d, (np.float64(ya) - gy) * cpm_d
        dsx = dsy = dth = F(0)
        th0 = tha
    gxb, gyb = F(gsx), F(gsy)                         # global: only for the validity checks
    sxb, syb = F(gsx - sh_x), F(gsy - sh_y)           # window-relative: what the model computes with
    n = len(r)
    one = np.ones(n, F)
    ends = [fma32(dsx * one, plan.rho_lo, gxb * one)[0], fma32(dsx * one, plan.rho_hi, gxb * one)[0],
            fma32(dsy * one, plan.rho_lo, gyb * one)[0], fma32(dsy * one, plan.rho_hi, gyb * one)[0]]
    lo, hi = min(ends), max(ends)
    ok = (lo >= 1.0 and hi <= plan.coord_hi and abs(dsx) <= plan.max_shift and abs(dsy) <= plan.max_shift
          and abs(th0) <= F(3.15) and abs(dth) <= F(3.15))
    if not ok:
        return np.zeros(n, np.int64), np.zeros(n, bool)
    sx = fma32(dsx * one, rho, sxb * one) if interp else sxb * one
    sy = fma32(dsy * one, rho, syb * one) if interp else syb * one
    thr = fma32(dth * one, rho, th0 * one) if interp else th0 * one
    a = (thr - th).astype(F)
    err = rng.uniform(-1.3e-6, 1.3e-6, (2, n))
    s = (np.sin(a.astype(np.float64)) + err[0]).astype(F)
    c = (np.cos(a.astype(np.float64)) + err[1]).astype(F)
    rc = (r * F(grid.cells_per_meter)).astype(F)
    px, py = (rc * c).astype(F), (rc * s).astype(F)
    ex, ey = (px + sx).astype(F), (py + sy).astype(F)
    with np.errstate(invalid="ignore", over="ignore"):
        bx = (ex + plan.magic).astype(F).view(np.int32)
        by = (ey + plan.magic).astype(F).view(np.int32)
        frac_ok = np.minimum((bx & plan.fmask).astype(np.uint32), (by & plan.fmask).astype(np.uint32)) != 0
        in_win = (np.abs((ex - plan.mid_x).astype(F)) < plan.half_x) & (np.abs((ey - plan.mid_y).astype(F)) < plan.half_y)
        cell_ok = frac_ok & in_win
        ax, ay = np.abs(px), np.abs(py)
        d1 = ((ax + ax).astype(F) - ay).astype(F)
        d2 = ((ay + ay).astype(F) - ax).astype(F)
        x2ok = ((ex + px).astype(F) >= plan.x2_lo_x) & ((ey + py).astype(F) >= plan.x2_lo_y)
        t_dir = np.where(x2ok, plan.t_dir, plan.t_dir_neg)
        dir_ok = np.minimum(np.abs(d1), np.abs(d2)) > t_dir
        outside = (np.abs((ex - plan.gmid_x).astype(F)) >= plan.ghalf_x) | (np.abs((ey - plan.gmid_y).astype(F)) >= plan.ghalf_y)
    offx = np.where(d1 > 0, np.where(np.signbit(px), -1, 1), 0)
    offy = np.where(d2 > 0, np.where(np.signbit(py), -1, 1), 0)
    cx = np.where(in_win, (bx >> plan.fb) - plan.mbk, 1) + plan.x0       # window cell -> global cell for the read
    cy = np.where(in_win, (by >> plan.fb) - plan.mbk, 1) + plan.y0

    def read(x, y):
        inside = (x >= 0) & (x < grid.width) & (y >= 0) & (y < grid.height)
        return np.where(inside, fast_cells[np.clip(y, 0, grid.height - 1), np.clip(x, 0, grid.width - 1)], 0).astype(np.int64)

    odds = read(cx, cy)
    o1 = read(cx - offx, cy - offy)          # toward the robot
    o2 = read(cx + offx, cy + offy)          # toward the doubled endpoint
    val_ok = cell_ok & ((odds > 0) | dir_ok)
    certain = val_ok | outside | (in_win & (odds == -1))
    v = np.where(odds > 0, 2 * odds, np.where(o1 > 0, o1, np.maximum(o2, 0)))
    return np.where(val_ok, v, 0), certain
