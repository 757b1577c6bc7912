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
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


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
        self.rho_abs, self.ang_room = F(rho_max * (1 + 1e-6)), F(9.5) - F(float(np.abs(thetas[valid]).max()))
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
          and abs(th0) <= F(3.15) and abs(dth) <= F(3.15)
          and fma32(plan.rho_abs, abs(dth), abs(th0)) <= plan.ang_room)
    if not ok:
        return np.zeros(n, np.int64), np.zeros(n, bool)
    sx = fma32(dsx * one, rho, sxb * one) if interp else sxb * one
    sy = fma32(dsy * one, rho, syb * one) if interp else syb * one
    thr = fma32(dth * one, rho, th0 * one) if interp else th0 * one
    thd = th.astype(np.float64)                                # score_table_kernel folds the beam angle into [-pi, pi]
    thf = np.where(thd > np.pi, thd - 2 * np.pi, np.where(thd < -np.pi, thd + 2 * np.pi, thd)).astype(F)
    a = (thr - thf).astype(F)
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


# ======================================================================================================================
# Model of the TABLE pass (botlab_b200/csrc/mcl_table.cuh: table_plan_kernel, make_tab_base, tab_eval, the K/T build).
# Same role as the model above: the design is checked on the CPU against the oracle's per-ray scores before (and
# independently of) the CUDA code.  Keep in step with mcl_table.cuh.
# ======================================================================================================================
TAB_FIXED = 129
TAB_B2 = F(0.59033447)
TAB_S = F(0.84697730)
TAB_C8 = F(1.27323954)
# sector s steps (ux, uy): 0 (+1,0) 1 (+1,+1) 2 (0,+1) 3 (-1,+1) 4 (-1,0) 5 (-1,-1) 6 (0,-1) 7 (+1,-1)
TAB_STEPS = [(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1)]


class TabPlan:
    """table_plan_kernel for a cloud: window = bounding box of poses and parents +- (longest ray + 6 cells), clipped to
    the grid plus four cells; coordinates are window-normalised ((c + kappa - 0.5) / (w - 1.5)) so that one saturating
    FFMA evaluates and clamps them."""

    def __init__(self, grid, cloud, ranges, thetas, ratios, min_range, smem_total=231424, smem_fixed=None):
        cpm = float(F(grid.cells_per_meter))
        valid = ranges > F(min_range)
        nb = int(valid.sum())
        self.ok = False
        if nb < 1 or nb > 2047 or not np.isfinite(ranges[valid]).all() or not np.isfinite(thetas[valid]).all():
            return
        max_range = float(ranges[valid].max())
        rc_max = max_range * cpm
        rho_lo, rho_hi = float(ratios[valid].min()), float(ratios[valid].max())
        rho_max = max(abs(rho_lo), abs(rho_hi))
        max_abs_theta = float(np.abs(thetas[valid]).max())
        if not (min_range * cpm >= 2.5 and max_abs_theta <= 6.3 and rho_lo >= -1.0 and rho_hi <= 2.0):
            return
        xs = np.concatenate([cloud["pose"]["x"], cloud["parent_pose"]["x"]])
        ys = np.concatenate([cloud["pose"]["y"], cloud["parent_pose"]["y"]])
        if not (np.isfinite(xs).all() and np.isfinite(ys).all()):
            return
        ox, oy = float(F(grid.origin_x)), float(F(grid.origin_y))
        reach = rc_max + 6.0
        cx0 = np.floor((float(xs.min()) - ox) * cpm - reach); cx1 = np.ceil((float(xs.max()) - ox) * cpm + reach)
        cy0 = np.floor((float(ys.min()) - oy) * cpm - reach); cy1 = np.ceil((float(ys.max()) - oy) * cpm + reach)
        if not (abs(cx0) < 1e6 and abs(cy0) < 1e6 and cx1 - cx0 < 8192 and cy1 - cy0 < 8192):
            return
        ux0, uy0, ux1, uy1 = int(cx0), int(cy0), int(cx1), int(cy1)
        x0, y0 = max(ux0, -4), max(uy0, -4)
        x1, y1 = min(ux1, grid.width + 3), min(uy1, grid.height + 3)
        tw, th_ = x1 - x0 + 1, y1 - y0 + 1
        if tw < 8 or th_ < 8:
            return
        pitch_k = (tw + 1) & ~1
        if ((pitch_k >> 1) & 1) == 0:
            pitch_k += 2
        k_bytes = (pitch_k * 2 * th_ + 15) & ~15
        if smem_fixed is None:
            nb4 = (nb + 3) & ~3
            smem_fixed = (nb4 * 36 + 32 * 192 * 2 + 1024 * 4 + 15) & ~15
        room = smem_total - smem_fixed - k_bytes - 64
        if room < 8 * (TAB_FIXED + 256):
            return
        cm_ = max(ux1, uy1) + 2.0
        if cm_ > 16000.0:
            return
        xm = cm_ / cpm + max(abs(ox), abs(oy))
        shift = 64.0
        ce = cm_ + rc_max
        e_ref = cpm * U * xm + 2 * U * ce + 2 * U * rc_max + rc_max * (20 * U + 1.2e-7) + 1e-9
        e_apx = 6 * U * max(tw, th_) + (1 + 2 * rho_max) * U * shift + 3 * U * rc_max + \
            rc_max * ((3.14159265358979 * (3 * rho_max + 2) + 9.5) * U + TRIG_ERR)
        self.eps = eps = 1.25 * (e_ref + e_apx) + 1e-6
        self.kappa = kappa = eps + 1e-5
        if kappa > 1.0 / 16.0:
            return
        self.ok = True
        self.x0, self.y0, self.w, self.h = x0, y0, tw, th_
        self.cap_entries = room // 8
        self.off = kappa - 0.5
        self.inv_sx, self.inv_sy = 1.0 / (tw - 1.5), 1.0 / (th_ - 1.5)
        self.mul_x, self.mul_y = (2 * tw - 3) << 8, (2 * th_ - 3) << 8
        self.bias_x, self.bias_y = 127 * tw - 191, 127 * th_ - 191
        self.frac_thr = int(np.ceil(2.0 * kappa * 4294967296.0))
        self.t3 = F(3.0 * (1.0 + 2.0 * eps) + 4.0 * U * rc_max + 1e-4)
        self.rho_lo, self.rho_hi, self.rho_abs = F(rho_lo), F(rho_hi), F(rho_max * (1 + 1e-6))
        self.max_shift, self.coord_hi = F(shift), F(cm_ - 1.0)
        self.reach = F(rc_max * (1 + 1e-6) + 4.0)
        self.ang_room = F(9.5) - min(F(max_abs_theta), F(3.1415928))     # the beam angles are folded into [-pi, pi]
        self.ulo_x, self.uhi_x = F(ux0 + 1.5 + float(self.reach)), F(ux1 + 1 - 1.5 - float(self.reach))
        self.ulo_y, self.uhi_y = F(uy0 + 1.5 + float(self.reach)), F(uy1 + 1 - 1.5 - float(self.reach))
        self.hmin_x, self.hmin_y = self.bias_x - x0, self.bias_y - y0
        x2_min = 3.0 * eps + 2.0 * kappa + 1e-3
        self.x2_lo_x = F((x2_min - x0 + self.off) * self.inv_sx * (1 + 1e-6) + 1e-7)
        self.x2_lo_y = F((x2_min - y0 + self.off) * self.inv_sy * (1 + 1e-6) + 1e-7)


def build_score_table(grid, plan):
    """The K tile (class per window cell) and the table T (8 sector scores per class) of score_table_kernel."""
    cells = grid.cells.astype(np.int64)
    H, W = cells.shape

    def raw(gx, gy):                                         # arrays of global cells; 0 outside the grid
        inside = (gx >= 0) & (gx < W) & (gy >= 0) & (gy < H)
        return np.where(inside, cells[np.clip(gy, 0, H - 1), np.clip(gx, 0, W - 1)], 0)

    gy, gx = np.meshgrid(np.arange(plan.y0, plan.y0 + plan.h), np.arange(plan.x0, plan.x0 + plan.w), indexing="ij")
    c = raw(gx, gy)
    near5 = np.zeros(c.shape, bool)
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            near5 |= raw(gx + dx, gy + dy) > 0
    nb = np.stack([raw(gx + ux, gy + uy) for ux, uy in TAB_STEPS])          # [8, h, w]
    any8 = (nb > 0).any(axis=0)
    far = (gx <= -4) | (gy <= -4) | (gx >= W + 3) | (gy >= H + 3)     # the truncated cell and its neighbours are outside
    neg = ~far & ((gx < 0) | (gy < 0))                                # truncation != floor: never certified
    gen = ~far & ~neg
    K = np.zeros(c.shape, np.int64)
    K[neg] = 1
    K[gen & (c <= 0) & near5 & ~any8] = 1
    K[gen & (c > 0)] = 1 + c[gen & (c > 0)]
    nt = gen & (c <= 0) & any8
    idx = np.flatnonzero(nt.ravel())
    T = np.zeros((TAB_FIXED + len(idx), 8), np.int64)
    for i in range(2, TAB_FIXED):
        T[i, :] = 2 * (i - 1)
    K.ravel()[idx] = TAB_FIXED + np.arange(len(idx))
    nbf = nb.reshape(8, -1)[:, idx]
    for s in range(8):
        o1, o2 = nbf[(s + 4) & 7], nbf[s]
        T[TAB_FIXED:, s] = np.where(o1 > 0, o1, np.maximum(o2, 0))
    return K, T, len(idx)


def table_pass(grid, plan, K, T, particle, ranges, thetas, ratios, min_range, rng, interp=True, force_edge=None):
    """One particle, all valid beams: (half_unit_scores, certain, edge).  rng perturbs the sine/cosine as in fast_pass."""
    valid = ranges > F(min_range)
    r, th, rho = ranges[valid], thetas[valid], ratios[valid].astype(F)
    n = len(r)
    gx, gy, cpm_d = np.float64(F(grid.origin_x)), np.float64(F(grid.origin_y)), np.float64(F(grid.cells_per_meter))
    xa, ya, tha = (F(particle["pose"][k]) for k in ("x", "y", "theta"))
    xb, yb, thb = (F(particle["parent_pose"][k]) for k in ("x", "y", "theta"))
    ddx = ddy = np.float64(0.0)
    if interp:
        gsx, gsy = (np.float64(xb) - gx) * cpm_d, (np.float64(yb) - gy) * cpm_d
        ddx, ddy = np.float64(F(xa - xb)) * cpm_d, np.float64(F(ya - yb)) * cpm_d
        d = np.float64(tha) - np.float64(thb)
        if abs(d) > np.pi:
            d += -2 * np.pi if d > 0 else 2 * np.pi
        th0, dth = thb, F(d)
    else:
        gsx, gsy = (np.float64(xa) - gx) * cpm_d, (np.float64(ya) - gy) * cpm_d
        dth = F(0)
        th0 = tha
    sxn, syn = F(((gsx - plan.x0) + plan.off) * plan.inv_sx), F(((gsy - plan.y0) + plan.off) * plan.inv_sy)
    dsxn, dsyn = F(ddx * plan.inv_sx), F(ddy * plan.inv_sy)
    gxb, gyb, dsx, dsy = F(gsx), F(gsy), F(ddx), F(ddy)
    xs = [fma32(dsx, plan.rho_lo, gxb), fma32(dsx, plan.rho_hi, gxb)]
    ys = [fma32(dsy, plan.rho_lo, gyb), fma32(dsy, plan.rho_hi, gyb)]
    xlo, xhi, ylo, yhi = min(xs), max(xs), min(ys), max(ys)
    lo, hi = min(xlo, ylo), max(xhi, yhi)
    in_uwin = xlo >= plan.ulo_x and xhi <= plan.uhi_x and ylo >= plan.ulo_y and yhi <= plan.uhi_y
    ok = (lo >= 1.0 and hi <= plan.coord_hi and abs(dsx) <= plan.max_shift and abs(dsy) <= plan.max_shift
          and abs(th0) <= F(3.15) and abs(dth) <= F(3.15)
          and fma32(plan.rho_abs, abs(dth), abs(th0)) <= plan.ang_room and in_uwin)
    if not ok:
        return np.zeros(n, np.int64), np.zeros(n, bool), -1
    edge = 0 if lo >= F(2.0) * plan.reach else (1 if lo >= plan.reach else 2)
    if force_edge is not None:
        edge = max(edge, force_edge)
    onev = np.ones(n, F)
    sxm = fma32(dsxn * onev, rho, sxn * onev) if interp else sxn * onev
    sym = fma32(dsyn * onev, rho, syn * onev) if interp else syn * onev
    thr = fma32(dth * onev, rho, th0 * onev) if interp else th0 * onev
    thd = th.astype(np.float64)                                # score_table_kernel folds the beam angle into [-pi, pi]
    thf = np.where(thd > np.pi, thd - 2 * np.pi, np.where(thd < -np.pi, thd + 2 * np.pi, thd)).astype(F)
    a = (thr - thf).astype(F)
    err = rng.uniform(-1.3e-6, 1.3e-6, (2, n))
    s = (np.sin(a.astype(np.float64)) + err[0]).astype(F)
    c = (np.cos(a.astype(np.float64)) + err[1]).astype(F)
    rc = (r * F(grid.cells_per_meter)).astype(F)
    rcx, rcy = (rc.astype(np.float64) * plan.inv_sx).astype(F), (rc.astype(np.float64) * plan.inv_sy).astype(F)
    nx, ny = np.clip(fma32(rcx, c, sxm), F(0), F(1)), np.clip(fma32(rcy, s, sym), F(0), F(1))      # fma.rn.sat.f32
    bxf, byf = (nx + F(1)).astype(F), (ny + F(1)).astype(F)
    wx = bxf.view(np.uint32).astype(np.uint64) * np.uint64(plan.mul_x)
    wy = byf.view(np.uint32).astype(np.uint64) * np.uint64(plan.mul_y)
    fx, hx = (wx & np.uint64(0xffffffff)).astype(np.int64), (wx >> np.uint64(32)).astype(np.int64)
    fy, hy = (wy & np.uint64(0xffffffff)).astype(np.int64), (wy >> np.uint64(32)).astype(np.int64)
    cell_ok = np.minimum(fx, fy) >= plan.frac_thr
    if edge >= 2:
        cell_ok = cell_ok & (hx >= plan.hmin_x) & (hy >= plan.hmin_y)
    cx, cy = hx - plan.bias_x, hy - plan.bias_y                 # window cells
    x2_pos = np.ones(n, bool)
    if edge >= 1:
        x2_pos = (fma32(rcx, c, nx) >= plan.x2_lo_x) & (fma32(rcy, s, ny) >= plan.x2_lo_y)
    assert (cx >= 0).all() and (cx < plan.w).all() and (cy >= 0).all() and (cy < plan.h).all(), "table pass read outside its window"
    Kv = K[cy, cx]
    r8 = fma32(a, TAB_C8, F(25165824.0) * onev)
    u8 = fma32(a, TAB_C8, -(r8 - F(25165824.0)).astype(F))
    wq = fma32(u8, TAB_S, (r8 - F(12582912.0)).astype(F))
    sector = wq.view(np.uint32).astype(np.int64) & 7
    v = T[Kv, sector]
    xq = np.float64(plan.t3) / (2.2360679 * rc.astype(np.float64) * (1.0 - 1e-6))
    with np.errstate(invalid="ignore"):
        d8 = np.where((xq >= 0.3) | ~(xq >= 0.0), 4.0, (np.arcsin(np.minimum(xq, 1.0)) + 2e-5) * 1.2732395447351628 * (1 + 1e-6))
    d8 = d8.astype(F) + F(1e-6)
    dir_ok = x2_pos & (np.abs((np.abs(u8) - TAB_B2).astype(F)) > d8)
    certain = (Kv == 0) | (cell_ok & ((Kv < TAB_FIXED) | dir_ok))
    return np.where(certain, v, 0), certain, edge


# ---------------------------------------------------------------------------------------------------------------------
# Beam culling of the score-table pass (mcl_table.cuh: table_bbox_kernel's heading range, table_plan_kernel's cull
# fields, table_cull_kernel): a beam is left out when every class-map cell its endpoints can fall into is class 0.
# ---------------------------------------------------------------------------------------------------------------------
class _Window:
    def __init__(self, x0, y0, w, h):
        self.x0, self.y0, self.w, self.h = x0, y0, w, h


def class_zero_map(grid, apron=4):
    """Boolean map over the grid plus the apron: True where the class map holds class 0 (index [gy + apron, gx + apron])."""
    K, _, _ = build_score_table(grid, _Window(-apron, -apron, grid.width + 2 * apron, grid.height + 2 * apron))
    return K == 0


def _fold(d):
    return np.where(np.abs(d) > np.pi, d - np.sign(d) * 2 * np.pi, d)


def cull_flags(grid, cloud, ranges, thetas, ratios, min_range, margin=4.5, hw_scale=1.0, apron=4):
    """Per valid beam: True = culled.  margin / hw_scale exist so that a test can show the rule is not vacuous."""
    valid = ranges > F(min_range)
    r, th = ranges[valid], thetas[valid]
    flags = np.zeros(len(r), bool)
    rho = ratios[valid]
    ta, tb = cloud["pose"]["theta"], cloud["parent_pose"]["theta"]
    if not (rho.min() >= 0.0 and rho.max() <= 1.0 and (np.abs(ta) <= F(3.15)).all() and (np.abs(tb) <= F(3.15)).all()):
        return flags
    th_ref = np.float64(tb[0])
    db = _fold(tb.astype(np.float64) - th_ref)
    da = db + _fold(ta.astype(np.float64) - tb.astype(np.float64))
    a0 = np.nextafter(F(min(db.min(), da.min())), F(-np.inf))          # rounded outward, as __double2float_rd / _ru
    a1 = np.nextafter(F(max(db.max(), da.max())), F(np.inf))
    if not (np.isfinite(a0) and np.isfinite(a1) and a1 - a0 < F(2.5)):
        return flags
    cpm = np.float64(F(grid.cells_per_meter))
    ox, oy = np.float64(F(grid.origin_x)), np.float64(F(grid.origin_y))
    xs = np.concatenate([cloud["pose"]["x"], cloud["parent_pose"]["x"]]).astype(np.float64)
    ys = np.concatenate([cloud["pose"]["y"], cloud["parent_pose"]["y"]]).astype(np.float64)
    bx0, bx1, by0, by1 = (xs.min() - ox) * cpm, (xs.max() - ox) * cpm, (ys.min() - oy) * cpm, (ys.max() - oy) * cpm
    ca0, ca1 = th_ref + np.float64(a0), th_ref + np.float64(a1)
    zero = class_zero_map(grid, apron)
    H, W = grid.height, grid.width
    cx, cy, hx, hy = 0.5 * (bx0 + bx1), 0.5 * (by0 + by1), 0.5 * (bx1 - bx0), 0.5 * (by1 - by0)
    for b in range(len(r)):
        rc = np.float64(F(r[b]) * F(grid.cells_per_meter))
        am = 0.5 * (ca0 + ca1) - np.float64(th[b])
        hw = (0.5 * (ca1 - ca0) * (1 + 1e-6) + 1e-6) * hw_scale
        cs, sn = np.cos(am), np.sin(am)
        bu, bv = hx * abs(cs) + hy * abs(sn), hx * abs(sn) + hy * abs(cs)
        u0, u1 = rc * np.cos(hw) * (1 - 1e-6) - bu - margin, rc * (1 + 1e-6) + bu + margin
        vm = rc * np.sin(hw) * (1 + 1e-6) + bv + margin
        ex = max(abs(u0), abs(u1)) * abs(cs) + vm * abs(sn)
        ey = max(abs(u0), abs(u1)) * abs(sn) + vm * abs(cs)
        x0, x1 = max(int(np.floor(cx - ex)), -apron), min(int(np.ceil(cx + ex)), W + apron - 1)
        y0, y1 = max(int(np.floor(cy - ey)), -apron), min(int(np.ceil(cy + ey)), H + apron - 1)
        if x1 < x0 or y1 < y0:
            flags[b] = True                                   # wholly beyond the apron: class 0 everywhere
            continue
        gy, gx = np.meshgrid(np.arange(y0, y1 + 1), np.arange(x0, x1 + 1), indexing="ij")
        dx, dy = gx - cx, gy - cy
        u, v = dx * cs + dy * sn, dy * cs - dx * sn
        inside = (u >= u0 - 0.01) & (u <= u1 + 0.01) & (np.abs(v) <= vm + 0.01)
        flags[b] = bool(zero[gy + apron, gx + apron][inside].all())
    return flags


# ---------------------------------------------------------------------------------------------------------------------
# 8-bit class tile of the score-table pass (mcl_table.cuh, WIDE): classes 0..128 as in K; a cell with a table entry
# holds 129 + its number within its 64-cell row segment, segments being aligned in (cell + bias_x), and the row's
# segment bases complete the entry's index.
# ---------------------------------------------------------------------------------------------------------------------
def tab_nseg(tw):
    bias = 127 * tw - 191
    return ((bias + tw - 1) >> 6) - (bias >> 6) + 1


def build_score_table_wide(K):
    """From the 16-bit tile K (build_score_table): the 8-bit tile, the segment bases [h, nseg] and, per entry of the
    wide numbering, the 16-bit tile's entry it must equal."""
    h, w = K.shape
    bias_x = 127 * w - 191
    nseg = tab_nseg(w)
    K8 = np.where(K < TAB_FIXED, K, 0).astype(np.int64)
    base = np.zeros((h, nseg), np.int64)
    entry_of = []                                    # wide entry index - TAB_FIXED -> K16 entry
    for ty in range(h):
        seg = ((np.arange(w) + bias_x) >> 6) - (bias_x >> 6)
        assert seg.min() >= 0 and seg.max() < nseg
        for sg in range(nseg):
            cols = np.flatnonzero((seg == sg) & (K[ty] >= TAB_FIXED))
            assert len(cols) <= 64
            base[ty, sg] = TAB_FIXED + len(entry_of)
            for i, tx in enumerate(cols):            # numbered in column order, as the ballots do
                K8[ty, tx] = TAB_FIXED + i
                entry_of.append(K[ty, tx])
    assert K8.max() < 256
    return K8, base, np.array(entry_of, np.int64), bias_x


def wide_entry(K8, base, bias_x, cy, cx):
    """Entry index of cell (cy, cx) as tab_eval<WIDE> forms it: K for the uniform classes, else segment base + K - 129."""
    k = K8[cy, cx]
    sg = ((cx + bias_x) >> 6) - (bias_x >> 6)
    return np.where(k < TAB_FIXED, k, base[cy, sg] + k - TAB_FIXED)
