"""CPU checks of the mathematical facts the separable visualizer kernel (csrc/visualizer_rows.cu) rests on, on the
oracle's own arithmetic — no GPU: (1) under the export camera the blur's centre tap is separable in (column, row);
(2) the strict-float32 tap table has the coincidences the tap grouping uses; (3) a bilinear tap equals the horizontal
lerp followed by the hinge-weight vertical interpolation; (4) vertical weights of taps sharing dx may be merged;
(5) horizontal lerps of taps sharing dy add up to one weighted sum over ni = floor(2*scale) + 3 texel columns."""
import numpy as np

from oracle import glsl_np as G

F = np.float32


def centre_taps(u, Wr, Hr, bg_size):
    """texel-space position of visualizer.frag's undisplaced tap (:16-18) for every fragment"""
    f = G.varyings(u, Wr, Hr)
    cam = G.get_camera(u, f)
    vol = F(u.extra["iAudioVolume"])
    z = F(F(F(0.95) + F(0.01)*np.sin(F(u.iTime))) - F(0.02)*vol) - F(0.03)
    uv = G.zoom(G.gluv2stuv(cam.gluv), np.full(cam.gluv.shape[:-1], z, F), G.vec(F(0.5), F(0.5)))
    uv = (uv + F(0.005)*G.vec(np.cos(F(u.iTime)*F(3.25135)), np.sin(F(u.iTime)*F(1.153469)))).astype(F)
    W, H = bg_size
    g = G.stuv2gluv(uv)*G.vec(F(H)/F(W), F(1))
    st = G.gluv2stuv(g.astype(F))
    return (st[..., 0]*F(W) - F(0.5)).astype(F), (st[..., 1]*F(H) - F(0.5)).astype(F)


def test_centre_tap_is_separable_under_the_export_camera():
    u = G.Uniforms(iTime=1.3, iResolution=(192, 108), iWantAspect=192/108, extra=dict(iAudioVolume=0.8, iAudioSTD=0.2))
    tx, ty = centre_taps(u, 384, 216, (960, 540))
    assert (tx == tx[:1, :]).all() and (ty == ty[:, :1]).all()          # exactly: x from the column, y from the row
    step_x, step_y = np.diff(tx[0]).mean(), np.diff(ty[:, 0]).mean()
    assert abs(step_x - step_y) < 1e-3 and 0.5 < step_x < 2.6            # one affine step per axis
    # a rotated basis breaks it (the launcher then picks the per-pixel kernel)
    c, s = np.cos(0.3), np.sin(0.3)
    u.iCameraRight, u.iCameraUpward = (c, s, 0.0), (-s, c, 0.0)
    tx, ty = centre_taps(u, 96, 54, (960, 540))
    assert not (tx == tx[:1, :]).all()


def blur_table():
    """dir*walk of visualizer.frag:26-27 with strict float32 counters → (rays, walks, 2)"""
    taps = []
    angle, step = F(0.0), F(G.TAU/F(8.0))
    while angle < G.TAU:
        ray, walk = [], F(F(1.0)/F(10.0))
        while walk <= F(1.001):
            ray.append((np.cos(angle)*walk, np.sin(angle)*walk))
            walk = F(walk + F(F(1.0)/F(10.0)))
        taps.append(ray)
        angle = F(angle + step)
    return np.array(taps, F)


def test_tap_table_coincidences_used_by_the_grouping():
    t = blur_table()
    assert t.shape == (9, 10, 2)                                          # float counters: 9 directions x 10 walks
    tol = 1e-6
    assert np.abs(t[8] - t[0]).max() < tol                                # the 9th direction is the first again
    assert np.abs(t[0, :, 1]).max() < tol and np.abs(t[4, :, 1]).max() < tol           # rays 0°, 180°: dy = 0
    assert np.abs(t[2, :, 0]).max() < tol and np.abs(t[6, :, 0]).max() < tol           # rays 90°, 270°: dx = 0
    assert np.abs(t[1, :, 1] - t[3, :, 1]).max() < tol and np.abs(t[5, :, 1] - t[7, :, 1]).max() < tol   # pairs share dy
    # at most 3.3 texels of reach on a 1080-row background: 0.01*0.3*1080
    assert abs(np.abs(t).max() - 1.0) < 1e-6


def test_hinge_weights_reproduce_the_bilinear_tap():
    rng = np.random.default_rng(0)
    tex = rng.integers(0, 256, (12, 16, 3)).astype(np.float64)
    for _ in range(200):
        px, py0 = rng.uniform(1, 13), rng.uniform(1, 7)
        ix, a = int(np.floor(px)), px - np.floor(px)
        # 8 fragment rows spanning < 2 texels share the 4 texel rows r0 .. r0+3
        rows = py0 + np.arange(8)*0.2325
        r0 = int(np.floor(rows.min()))
        H = tex[r0:r0 + 4, ix] + a*(tex[r0:r0 + 4, ix + 1] - tex[r0:r0 + 4, ix])        # horizontal lerp per texel row
        D = np.diff(H, axis=0)
        for py in rows:
            t = py - r0
            hinge = H[0] + sum(D[r]*np.clip(t - r, 0, 1) for r in range(3))
            iy, b = int(np.floor(py)), py - np.floor(py)
            top = tex[iy, ix]*(1 - a) + tex[iy, ix + 1]*a
            bot = tex[iy + 1, ix]*(1 - a) + tex[iy + 1, ix + 1]*a
            assert np.allclose(hinge, top*(1 - b) + bot*b, atol=1e-9)


def test_vertical_weights_of_taps_sharing_dx_merge():
    """sum_k V_k(H) = sum_rows H_row * sum_k hat(py_k - row): the 21 dx = 0 taps cost one pass over the texel rows"""
    rng = np.random.default_rng(1)
    H = rng.uniform(0, 255, (16, 3))                                       # one horizontally interpolated column of texel rows
    hat = lambda d: np.maximum(0.0, 1.0 - np.abs(d))
    dys = np.concatenate([np.arange(1, 11)*0.324, -np.arange(1, 11)*0.324, [0.0]])
    py = 7.3 + dys
    direct = sum(H[int(np.floor(p))]*(1 - (p - np.floor(p))) + H[int(np.floor(p)) + 1]*(p - np.floor(p)) for p in py)
    merged = sum(H[row]*hat(py - row).sum() for row in range(16))
    assert np.allclose(direct, merged, atol=1e-9)


def test_horizontal_weights_of_taps_sharing_dy_merge():
    """Phase 2 of the rows kernel: the 20 taps of rays 0° (weighted twice: it also stands for the 9th direction) and 180°
    read the same texel rows, so sum_k w_k * lerp(T, px_k) = sum_i T[i] * sum_k w_k * hat(px_k - i) over the
    ni = floor(2*scale*1.0001) + 3 texel columns starting at floor(cx - scale*1.0001) — never more, whatever the phase"""
    rng = np.random.default_rng(2)
    t = blur_table()
    hat = lambda d: np.maximum(0.0, 1.0 - np.abs(d))
    dx = np.concatenate([t[0, :, 0], t[4, :, 0]]).astype(np.float64)          # ray 0 (10 walks), ray 180
    weight = np.concatenate([np.full(10, 2.0), np.full(10, 1.0)])
    for scale in (0.0, 0.05, 0.53, 1.0, 2.499, 2.5, 3.24):
        ni = int(2.0*scale*1.0001) + 3
        for _ in range(200):
            T = rng.integers(0, 256, (40, 3)).astype(np.float64)               # one texel row
            cx = rng.uniform(6.0, 30.0)
            px = cx + dx*scale
            i, a = np.floor(px).astype(int), px - np.floor(px)
            direct = (weight[:, None]*(T[i] + a[:, None]*(T[i + 1] - T[i]))).sum(axis=0)
            ix0 = int(np.floor(cx - scale*1.0001))
            wx = np.array([(weight*hat(px - (ix0 + k))).sum() for k in range(ni)])
            assert np.allclose(direct, (wx[:, None]*T[ix0:ix0 + ni]).sum(axis=0), atol=1e-9)
            assert abs(wx.sum() - 30.0) < 1e-9                                 # every tap's two weights are inside the ni columns
