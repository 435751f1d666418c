"""
TEST INFRASTRUCTURE ONLY (oracle, kind "port"). Never imported by the product path; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it.

numpy float32 restatement of the reference's GLSL render path, one array lane per shaded sample:

    varyings                      shaderflow/resources/shaders/vertex/default.glsl:1-17, shader.py:127-128
    std-lib helpers               shaderflow/resources/shaders/include/shaderflow.glsl (line refs inline)
    camera                        shaderflow/resources/shaders/include/camera.glsl:55-155
    SSAA downsample               shaderflow/resources/shaders/fragment/final.glsl:3-33
    scenes                        fragment/default.glsl, examples/basic/shaders/{visualizer,bars,waveform,
                                  shadertoy,raymarch,multipass,motionblur}.frag, life/{simulation,visuals}.glsl,
                                  the inline GLSL of examples/basic/demo.py (MultiShader, Dynamics, Audio),
                                  examples/fractals/shaders/{mandelbrot,tetration}.frag
    texture formats / filters     shaderflow/texture.py:28-38,104-137,175-182,327-338
    render / readback conventions shaderflow/shader.py:367-405, scene.py:185-194, exporting.py:94-103,165-174

PINNED TO THE REFERENCE'S SHADER TEXT (round 2). No OpenGL implementation exists in the build container or on
the GPU box, so the reference cannot render here. Instead the GLSL the reference assembles and hands to the
driver is captured from the reference's own Python (`oracle/ref_scene.py`) and executed by a mechanical GLSL
evaluator (`oracle/glsl_exec.py`); the results are committed as `tests/golden/glsl_*.npz`
(`tests/golden/make_golden_glsl.py`) and `tests/test_oracle_glsl.py` holds every function below to them at
≤ 1e-6 per float channel — 16 scene programs incl. stereo / equirectangular / rotated cameras, final.glsl over
9 (ssaa, subsample) geometries, and four 8-row bands of the benchmarked 7680×4320 target. What is still a
restatement of a SPECIFICATION (OpenGL 3.3 / GLSL 3.30), shared with the evaluator, is fixed as follows:
  * fragment (i, j) of a Wr×Hr target, j=0 at the BOTTOM, is shaded at its centre; every varying is the
    exact affine interpolation of the vertex-shader outputs at the quad corners, rounded once to float32;
  * texture(): normalised coords, LINEAR = weights from frac(u·W−0.5) on texels floor(u·W−0.5)+{0,1},
    NEAREST = floor(u·W); REPEAT wraps texel indices modulo the size, CLAMP_TO_EDGE clamps them;
    unorm8 texels decode as c/255; a 3-component image reads alpha=1;
  * colour stores to an 8-bit target clamp to [0,1] and round-half-even (rint(c·255)); NaN stores 0;
  * float arithmetic is strict IEEE float32, one rounding per operation, no FMA contraction
    (visualizer.frag:26 → 9 directions, App. D-11);
  * clamp/min/max drop NaNs the way fminf/fmaxf do; `int(floor(NaN))` in hsv2rgb takes `default:`.
All arithmetic is float32 (numpy keeps float32 when combined with Python scalars under NEP 50).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

F = np.float32
PI, TAU = F(3.1415926535897932), F(6.2831853071795864)
SQRT3 = F(1.7320508075688772)

# ---------------------------------------------------------------------------------------------- #
# GLSL builtins on arrays (vectors keep their components on the last axis)

def vec(*parts):
    parts = np.broadcast_arrays(*[np.asarray(p, F) for p in parts])
    return np.stack(parts, axis=-1)

def fclamp(x, lo, hi):
    return np.fmin(np.fmax(np.asarray(x, F), F(lo)), F(hi))

def mix(a, b, t):
    t = np.asarray(t, F)
    return (a*(F(1) - t) + b*t).astype(F)

def smoothstep(e0, e1, x):
    t = fclamp((np.asarray(x, F) - F(e0))/(F(e1) - F(e0)), 0, 1)
    return (t*t*(F(3) - F(2)*t)).astype(F)

def gmod(x, y):
    x = np.asarray(x, F); y = F(y)
    return (x - y*np.floor(x/y)).astype(F)

def dot(a, b):
    p = (a*b).astype(F)
    s = p[..., 0]
    for i in range(1, p.shape[-1]):
        s = (s + p[..., i]).astype(F)
    return s

def length(v):
    return np.sqrt(dot(v, v)).astype(F)

def normalize(v):
    return (v/length(v)[..., None]).astype(F)

def gpow(x, y):
    with np.errstate(all="ignore"):
        return np.power(np.asarray(x, F), np.asarray(y, F)).astype(F)

# ---------------------------------------------------------------------------------------------- #
# shaderflow.glsl helpers (line numbers of include/shaderflow.glsl)

def gluv2stuv(g):                    # :100
    return ((g + F(1))/F(2)).astype(F)

def stuv2gluv(s):                    # :96
    return (s*F(2) - F(1)).astype(F)

def zoom(uv, z, anchor):             # :361-363
    z = np.asarray(z, F)
    return ((uv - anchor)*(z*z)[..., None] + anchor).astype(F)

def rotate2d_apply(angle, v):
    """mat2(c,-s,s,c) is column-major (:75-77): columns (c,-s) and (s,c); M·v = (c·x+s·y, −s·x+c·y)"""
    c, s = np.cos(F(angle)), np.sin(F(angle))
    return vec(c*v[..., 0] + s*v[..., 1], (-s)*v[..., 0] + c*v[..., 1])

def atan1n(p):                       # :378-380
    return (np.arctan2(p[..., 1], p[..., 0])/PI).astype(F)

def atan2_pos(y, x):                 # :382-388
    y = np.asarray(y, F); x = np.asarray(x, F)
    return np.where(y < 0, TAU - np.arctan2(-y, x), np.arctan2(y, x)).astype(F)

def palette(t, A, B, C, D):          # :212-220
    t = np.asarray(t, F)[..., None]
    A, B, C, D = (np.asarray(v, F) for v in (A, B, C, D))
    return np.where(t < F(0.25), mix(A, B, t*F(4)),
           np.where(t < F(0.5), mix(B, C, (t - F(0.25))*F(4)), mix(C, D, (t - F(0.5))*F(4)))).astype(F)

MAGMA = ((0.01060815, 0.01808215, 0.10018654), (0.38092887, 0.12061482, 0.32506528),
         (0.79650140, 0.10506637, 0.31063031), (0.95922872, 0.53307513, 0.37488950))   # :222-226

def palette_magma(t):
    return palette(t, *MAGMA)

def hsv2rgb(h, s, v):                # :406-425
    h = np.asarray(h, F); s = np.asarray(s, F); v = np.asarray(v, F)
    h = gmod(h, TAU)
    c = (v*s).astype(F)
    x = (c*(F(1) - np.abs(gmod(h/(PI/F(3)), 2) - F(1)))).astype(F)
    m = (v - c).astype(F)
    with np.errstate(invalid="ignore"):
        sector = np.floor(F(6)*(h/(F(2)*PI)))
    sector = np.where(np.isnan(sector), F(-1), sector)
    z = np.zeros_like(c + x)
    table = [vec(c, x, z), vec(x, c, z), vec(z, c, x), vec(z, x, c), vec(x, z, c), vec(c, z, x)]
    rgb = np.zeros(np.broadcast(c, x).shape + (3,), F)
    for i, t in enumerate(table):
        rgb = np.where((sector == i)[..., None], t, rgb)
    return (rgb + m[..., None]).astype(F)

def sdBox(origin, point, size):      # :285-288
    d = (np.abs(origin - np.asarray(point, F)) - np.asarray(size, F)/F(2)).astype(F)
    inner = np.fmin(np.fmax(d[..., 0], np.fmax(d[..., 1], d[..., 2])), F(0))
    return (inner + length(np.fmax(d, F(0)))).astype(F)

# ---------------------------------------------------------------------------------------------- #
# Textures

@dataclass
class Texture:
    """A GL 2D texture: `data` (H, W, C) with row 0 = v=0 (the BOTTOM row; from_numpy flips images,
    texture.py:334). uint8 data is unorm8, float32 data is read as is."""
    data: np.ndarray
    linear: bool = True
    repeat_x: bool = True
    repeat_y: bool = True

    @property
    def size(self):
        return (self.data.shape[1], self.data.shape[0])

    def texels(self, ix, iy):
        H, W, C = self.data.shape
        ix = np.mod(ix, W) if self.repeat_x else np.clip(ix, 0, W - 1)
        iy = np.mod(iy, H) if self.repeat_y else np.clip(iy, 0, H - 1)
        t = self.data[iy, ix]
        t = (t.astype(F)/F(255)) if self.data.dtype == np.uint8 else t.astype(F)
        if C < 4:
            pad = [np.zeros(t.shape[:-1] + (1,), F)]*(3 - C) + [np.ones(t.shape[:-1] + (1,), F)]
            t = np.concatenate([t] + pad, axis=-1)
        return t

    def sample(self, uv, offset=(0, 0)):
        """texture(sampler, uv) → (..., 4) float32; `offset` (textureOffset, GL 3.3 §3.8.7) is a whole number of
        texels added to the texel coordinates before the wrap mode applies"""
        uv = np.asarray(uv, F)
        W, H = self.size
        ox, oy = (np.asarray(o, np.int64) for o in offset)
        u = (uv[..., 0]*F(W)).astype(F)
        v = (uv[..., 1]*F(H)).astype(F)
        if not self.linear:
            return self.texels(np.floor(u).astype(np.int64) + ox, np.floor(v).astype(np.int64) + oy)
        ub, vb = (u - F(0.5)).astype(F), (v - F(0.5)).astype(F)
        fx, fy = np.floor(ub), np.floor(vb)
        a, b = (ub - fx).astype(F)[..., None], (vb - fy).astype(F)[..., None]
        i0, j0 = fx.astype(np.int64) + ox, fy.astype(np.int64) + oy
        t00, t10 = self.texels(i0, j0), self.texels(i0 + 1, j0)
        t01, t11 = self.texels(i0, j0 + 1), self.texels(i0 + 1, j0 + 1)
        top = (t00*(F(1) - a) + t10*a).astype(F)
        bot = (t01*(F(1) - a) + t11*a).astype(F)
        return (top*(F(1) - b) + bot*b).astype(F)


def gtexture(tex: Texture, gluv):    # shaderflow.glsl:165-169
    W, H = tex.size
    scale = vec(F(H)/F(W), F(1))
    return tex.sample(gluv2stuv((gluv*scale).astype(F)))

def stexture(tex: Texture, stuv):    # :202-204
    return gtexture(tex, stuv2gluv(stuv))

# ---------------------------------------------------------------------------------------------- #
# Uniforms, varyings, camera

@dataclass
class Uniforms:
    """Values as the GL program would hold them (float32). Defaults = a freshly built scene in an
    export: scene.py:687-703, camera.py:146-201 (position 0, zoom 1, isometric 0, focal 1, mode 2D)"""
    iTime: float = 0.0
    iTau: float = 0.0
    iDuration: float = 10.0
    iResolution: tuple = (1920, 1080)
    iWantAspect: float = 16/9
    iQuality: float = 0.5
    iSSAA: float = 1.0
    iFramerate: float = 60.0
    iFrame: int = 0
    iLayer: int = 0
    iCameraMode: int = 1
    iCameraProjection: int = 0
    iCameraPosition: tuple = (0.0, 0.0, 0.0)
    iCameraRight: tuple = (1.0, 0.0, 0.0)
    iCameraUpward: tuple = (0.0, 1.0, 0.0)
    iCameraForward: tuple = (0.0, 0.0, 1.0)
    iCameraZenith: tuple = (0.0, 1.0, 0.0)
    iCameraZoom: float = 1.0
    iCameraIsometric: float = 0.0
    iCameraFocalLength: float = 1.0
    iCameraOrbital: float = 0.0
    iCameraDolly: float = 0.0
    iCameraSeparation: float = 0.05
    extra: dict = field(default_factory=dict)     # module uniforms: iAudioVolume, iAudioSTD, ...

    def __getitem__(self, name):
        return self.extra[name]

    @property
    def aspect(self):                # #define iAspectRatio (float(iResolution.x)/iResolution.y)
        return F(F(self.iResolution[0])/F(self.iResolution[1]))


@dataclass
class Frag:
    agluv: np.ndarray; gluv: np.ndarray; astuv: np.ndarray; stuv: np.ndarray
    stxy: np.ndarray; glxy: np.ndarray


def varyings(u: Uniforms, Wr: int, Hr: int, rows: slice | None = None) -> Frag:
    """vertex/default.glsl:8-16 evaluated at the quad corners (±1, shader.py:127-128), interpolated
    affinely to the fragment centres in float64, rounded once to float32"""
    i = (np.arange(Wr, dtype=np.float64) + 0.5)/Wr
    j = (np.arange(Hr, dtype=np.float64) + 0.5)/Hr
    if rows is not None:
        j = j[rows]
    tx, ty = np.meshgrid(i, j)
    W, H = float(F(u.iResolution[0])), float(F(u.iResolution[1]))
    a = float(u.aspect)
    def lerp2(x0, x1, y0, y1):
        return np.stack([x0 + (x1 - x0)*tx, y0 + (y1 - y0)*ty], -1).astype(F)
    s0, s1 = float((F(-a) + F(1))/F(2)), float((F(a) + F(1))/F(2))
    return Frag(
        agluv=lerp2(-1.0, 1.0, -1.0, 1.0),
        gluv =lerp2(float(F(-1)*F(a)), float(F(1)*F(a)), -1.0, 1.0),
        astuv=lerp2(0.0, 1.0, 0.0, 1.0),
        stuv =lerp2(s0, s1, 0.0, 1.0),
        stxy =lerp2(1.0, float(F(W)*F(1) + F(1)), 1.0, float(F(H)*F(1) + F(1))),
        glxy =lerp2(float(F(1) - F(W)/F(2)), float((F(W) + F(1)) - F(W)/F(2)),
                    float(F(1) - F(H)/F(2)), float((F(H) + F(1)) - F(H)/F(2))),
    )


@dataclass
class Camera:
    gluv: np.ndarray; agluv: np.ndarray; stuv: np.ndarray; astuv: np.ndarray
    origin: np.ndarray; target: np.ndarray; out_of_bounds: np.ndarray


def get_camera(u: Uniforms, f: Frag) -> Camera:
    """camera.glsl:55-155 (GetCamera → CameraProject → CameraRay2D)"""
    v3 = lambda t: np.asarray(t, F)
    pos, right, up, fwd = v3(u.iCameraPosition), v3(u.iCameraRight), v3(u.iCameraUpward), v3(u.iCameraForward)
    back = (fwd*F(-1)).astype(F)
    zoom_, iso, focal = F(u.iCameraZoom), F(u.iCameraIsometric), F(u.iCameraFocalLength)
    orbital, dolly = F(u.iCameraOrbital), F(u.iCameraDolly)
    aspect = u.aspect

    def rectangle(g, size):          # :55-57
        return (F(size)*(g[..., 0:1]*right + g[..., 1:2]*up)).astype(F)

    def ray_origin(p, g):            # :59-64
        return (((p + rectangle(g, zoom_*iso)) + back*orbital) + back*dolly).astype(F)

    def ray_target(p, g):            # :66-71
        return (((p + rectangle(g, zoom_)) + back*orbital) + fwd*focal).astype(F)

    if u.iCameraProjection == 0:
        origin, target = ray_origin(pos, f.gluv), ray_target(pos, f.gluv)
    elif u.iCameraProjection == 1:   # :101-110
        sgn = np.sign(f.agluv[..., 0:1]).astype(F)
        g = (f.gluv - sgn*vec(aspect/F(2), F(0))).astype(F)
        p = (pos + (sgn*F(u.iCameraSeparation))*right).astype(F)
        origin, target = ray_origin(p, g), ray_target(p, g)
    else:                            # :113-127
        def rotate3d(vector, axis, angle):        # shaderflow.glsl:82-84
            c, s = np.cos(angle)[..., None], np.sin(angle)[..., None]
            return (mix(dot(axis, vector)[..., None]*axis, vector, c) + np.cross(axis, vector)*s).astype(F)
        inc = (zoom_*(PI*f.agluv[..., 1]/F(2))).astype(F)
        azi = (zoom_*(PI*f.agluv[..., 0]/F(1))).astype(F)
        t = np.broadcast_to(fwd, f.gluv.shape[:-1] + (3,)).astype(F)
        t = rotate3d(t, right, -inc)
        t = rotate3d(t, up, azi)
        origin = np.broadcast_to(pos, t.shape).astype(F)
        target = (pos + t).astype(F)

    # CameraRay2D :73-91, plane through (0,0,1) with normal (0,0,1)
    pp, pn = v3((0, 0, 1)), v3((0, 0, 1))
    num = dot(pp - origin, pn)
    den = dot(target - origin, pn)
    with np.errstate(all="ignore"):
        t = (num/den).astype(F)
    oob = (t < 0) | (np.abs(f.gluv[..., 0]) > F(u.iWantAspect))
    hit = (origin + t[..., None]*(target - origin)).astype(F)
    g = hit[..., :2]
    ag = (g/vec(aspect, F(1))).astype(F)
    return Camera(gluv=g, agluv=ag, stuv=gluv2stuv(g), astuv=gluv2stuv(ag),
                  origin=origin, target=target, out_of_bounds=oob)

# ---------------------------------------------------------------------------------------------- #
# Scene fragments: (uniforms, varyings, textures) → fragColor (..., 4) float32

def frag_default(u: Uniforms, f: Frag, tex: dict):
    """fragment/default.glsl:10-48 (scene `Basic`)"""
    cam = get_camera(u, f)
    uv = cam.gluv
    ang = atan2_pos(uv[..., 1], uv[..., 0])
    color = (F(0.3) + hsv2rgb(ang + (F(2)*TAU*F(u.iTau)) - (PI/F(4)), 1, 1)).astype(F)
    circle = (F(1.333)*length(uv) - F(1.0)).astype(F)
    with np.errstate(all="ignore"):
        width = (F(2)*np.abs(F(1)/(circle*circle))*F(1e-4)).astype(F)
    chk = gmod(np.floor(uv[..., 0]*F(8.0)/F(2)) + np.floor(uv[..., 1]*F(8.0)/F(2)), 2.0) > F(0.5)
    grid = np.where(chk, F(0.22), F(0.20)).astype(F)
    rgb = np.where(circle < 0, F(0.18), grid).astype(F)[..., None].repeat(3, -1)
    with np.errstate(all="ignore"):
        rgb = (rgb + width[..., None]*color).astype(F)
    away = (f.astuv*(F(1.0) - f.astuv[..., ::-1])).astype(F)
    lin = (F(50)*(away[..., 0]*away[..., 1])).astype(F)
    rgb = (rgb*fclamp(gpow(lin, 0.1), 0.0, 1.0)[..., None]).astype(F)
    rgb = np.where(cam.out_of_bounds[..., None], F(0.15), rgb)
    return np.concatenate([rgb, np.ones(rgb.shape[:-1] + (1,), F)], -1).astype(F)


def frag_shadertoy(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/shadertoy.frag:62-66"""
    s = f.stuv
    xyx = vec(s[..., 0], s[..., 1], s[..., 0])
    col = (F(0.5) + F(0.5)*np.cos(F(u.iTime) + xyx + np.asarray((0, 2, 4), F))).astype(F)
    return np.concatenate([col, np.ones(col.shape[:-1] + (1,), F)], -1)


def visualizer_directions():
    """visualizer.frag:26-27 in strict float32: the (angle, walk) pairs the two float loops visit"""
    angles, a, step = [], F(0), F(TAU/F(8))
    while a < TAU:
        angles.append(a); a = F(a + step)
    walks, w, ws = [], F(F(1.0)/F(10)), F(F(1.0)/F(10))
    while w <= F(1.001):
        walks.append(w); w = F(w + ws)
    return angles, walks


def frag_visualizer(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/visualizer.frag:6-74"""
    cam = get_camera(u, f)
    uv = cam.gluv
    space = (np.asarray((1, 11, 26), F)/F(255)).astype(F)
    vol, std, time = F(u["iAudioVolume"]), F(u["iAudioSTD"]), F(u.iTime)
    bg = tex["background"]

    zf = F(F(F(F(0.95) + F(0.01)*np.sin(time)) - F(0.02)*vol) - F(0.03))
    buv = zoom(gluv2stuv(uv), zf, vec(F(0.5), F(0.5)))
    buv = (buv + F(0.005)*vec(np.cos(time*F(3.25135)), np.sin(time*F(1.153469)))).astype(F)
    color = stexture(bg, buv)

    intensity = F(F(0.01)*fclamp(gpow(vol, 2.5), 0, 0.3))
    acc = color.copy()
    angles, walks = visualizer_directions()
    for a in angles:
        d = vec(np.cos(a), np.sin(a))
        for w in walks:
            disp = ((d*w)*intensity).astype(F)
            acc = (acc + stexture(bg, (buv + disp).astype(F))).astype(F)
    color = (acc/F(F(10)*F(8))).astype(F)

    color = (color*(F(1) + F(5)*std*gpow(fclamp(length(f.agluv) - F(0.3), 0, 1), 6))[..., None]).astype(F)

    muv = rotate2d_apply(-PI/F(2), uv)
    muv = (muv*F(F(1) - F(0.4)*gpow(np.abs(vol), 0.5))).astype(F)
    radius = F(0.17)
    circle = np.abs(atan1n(muv))
    spec = tex["iSpectrogram"].sample(vec(np.zeros_like(circle), circle))[..., :2]
    with np.errstate(invalid="ignore"):
        freq = np.sqrt(spec/F(1000)).astype(F)
    freq = (freq*(F(0.05) + F(3)*smoothstep(0, 2, circle))[..., None]).astype(F)

    rgb = color[..., :3]
    lm = length(muv)
    bar = np.where(muv[..., 1] < 0, freq[..., 0], freq[..., 1]).astype(F)
    r = (radius + F(0.5)*bar).astype(F)
    inner = (rgb*F(0.5)).astype(F)
    on_bar = mix(rgb, np.ones(3, F), smoothstep(0, 1, F(0.5) + bar)[..., None])
    with np.errstate(invalid="ignore"):
        outside = (rgb*gpow((lm - r)*F(0.5), 0.05)[..., None]).astype(F)
    rgb = np.where((lm < radius)[..., None], inner, np.where((lm < r)[..., None], on_bar, outside)).astype(F)

    rgb = mix(rgb, space, smoothstep(0, 1, length(uv)/F(20))[..., None])

    vig = (f.astuv*(F(1) - f.astuv[..., ::-1])).astype(F)
    rgb = (rgb*gpow(vig[..., 0]*vig[..., 1]*F(20), F(0.1) + F(0.15)*vol)[..., None]).astype(F)
    out = np.concatenate([rgb, np.ones(rgb.shape[:-1] + (1,), F)], -1)

    wave = (F(0.2)*tex["iWaveform"].sample(vec(f.astuv[..., 0], np.zeros_like(f.astuv[..., 0])))[..., :2]).astype(F)
    out = np.where(((F(1) - f.gluv[..., 1]) < wave[..., 0])[..., None], (out*F(0.8)).astype(F), out)
    out = np.where(((F(1) + f.gluv[..., 1]) < wave[..., 1])[..., None], (out*F(0.8)).astype(F), out)

    oob = np.concatenate([np.broadcast_to(space, rgb.shape), np.zeros(rgb.shape[:-1] + (1,), F)], -1)
    return np.where(cam.out_of_bounds[..., None], oob, out).astype(F)


def frag_bars(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/bars.frag:5-22"""
    a = f.astuv
    with np.errstate(invalid="ignore"):
        inten = (np.sqrt(tex["iSpectrogram"].sample(a[..., ::-1])[..., :2])/F(120)).astype(F)
    y = a[..., 1]
    rgb = np.zeros(a.shape[:-1] + (3,), F)
    rgb[..., 0] += np.where(y < inten[..., 0], F(1), F(0))
    rgb[..., 1] += np.where(y < inten[..., 1], F(1), F(0))
    rgb[..., 2] += np.where(y < ((inten[..., 1] + inten[..., 0])/F(2.0)).astype(F), F(1), F(0))
    rgb[..., 2] = (rgb[..., 2] + (F(0.4)*(inten[..., 0] + inten[..., 1]))*(F(1.0) - y)).astype(F)
    return np.concatenate([rgb, np.ones(rgb.shape[:-1] + (1,), F)], -1)


def frag_waveform(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/waveform.frag:5-19"""
    wave = tex["iWaveform"].sample(vec(f.astuv[..., 0], np.zeros_like(f.astuv[..., 0])))[..., :2]
    ay = np.abs(f.gluv[..., 1])
    out = np.empty(ay.shape + (4,), F)
    out[..., :3] = F(0.2); out[..., 3] = 1
    out[..., 0] = np.where(ay < wave[..., 0], F(1), out[..., 0])
    out[..., 1] = np.where(ay < wave[..., 1], F(1), out[..., 1])
    out[..., 2] = np.where(ay < ((wave[..., 0] + wave[..., 1])/F(2)).astype(F), F(1), out[..., 2])
    return out


def frag_mandelbrot(u: Uniforms, f: Frag, tex: dict):
    """examples/fractals/shaders/mandelbrot.frag:10-31"""
    cam = get_camera(u, f)
    z = (cam.gluv - vec(F(0.5), F(0.0))).astype(F)
    c = z.copy()
    quality = int(F(1000.0)*F(u.iQuality))
    it = np.zeros(z.shape[:-1], np.int32)
    alive = np.ones(z.shape[:-1], bool)
    for _ in range(quality):
        with np.errstate(all="ignore"):
            alive &= ~(length(z) > F(3.0))
            if not alive.any():
                break
            zx, zy = z[..., 0], z[..., 1]
            nz = vec((zx*zx - zy*zy).astype(F) + c[..., 0], (zx*zy + zy*zx).astype(F) + c[..., 1])
        z = np.where(alive[..., None], nz, z)
        it += alive
    t = gpow(F(1) - it.astype(F)/F(quality), 20)
    rgb = np.where(cam.out_of_bounds[..., None], palette_magma(np.zeros_like(t)), palette_magma(t))
    return np.concatenate([rgb, np.ones(rgb.shape[:-1] + (1,), F)], -1).astype(F)


def frag_tetration(u: Uniforms, f: Frag, tex: dict):
    """examples/fractals/shaders/tetration.frag:28-55"""
    cam = get_camera(u, f)
    cx, cy = cam.gluv[..., 0], cam.gluv[..., 1]
    cr = np.sqrt((cx*cx + cy*cy).astype(F)).astype(F)
    ct = np.arctan2(cy, cx).astype(F)
    zx, zy, zr = cx.copy(), cy.copy(), cr.copy()
    MAX_STEPS = 67
    it = np.zeros(cx.shape, np.int32)
    alive = np.ones(cx.shape, bool)
    with np.errstate(all="ignore"):
        for _ in range(MAX_STEPS):
            r = (gpow(cr, zx)*np.exp(-zy*ct)).astype(F)
            t = (zy*np.log(cr) + (zx*ct)).astype(F)
            nx, ny = (r*np.cos(t)).astype(F), (r*np.sin(t)).astype(F)
            zx, zy, zr = np.where(alive, nx, zx), np.where(alive, ny, zy), np.where(alive, r, zr)
            alive = alive & ~(zr > F(100.0))     # break happens after the update, before it++
            it += alive
            if not alive.any():
                break
    k = (it // MAX_STEPS).astype(F)              # tetration.frag:50 integer division
    with np.errstate(all="ignore"):
        theta = (atan2_pos(zy, zx)/TAU).astype(F)
    col = hsv2rgb(theta, F(1.0), k)
    return np.concatenate([col, np.ones(col.shape[:-1] + (1,), F)], -1).astype(F)


def frag_raymarch(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/raymarch.frag:33-60"""
    cam = get_camera(u, f)
    origin = cam.origin
    forward = normalize((cam.target - origin).astype(F))
    traveled = np.zeros(origin.shape[:-1], F)
    steps = np.zeros(origin.shape[:-1], np.int32)
    alive = np.ones(origin.shape[:-1], bool)
    for _ in range(100):
        point = (origin + forward*traveled[..., None]).astype(F)
        sdf = np.full(traveled.shape, F(2)*F(100.0), F)
        for i in range(2, 8):
            sdf = np.fmin(sdf, sdBox(point, (0, 0, i), (i - 1,)*3))
        traveled = np.where(alive, (traveled + sdf).astype(F), traveled)
        alive = alive & ~((sdf < F(0.001)) | (sdf > F(100.0)))
        steps += alive
        if not alive.any():
            break
    col = (F(1) - np.sqrt(steps.astype(F))*F(0.1)).astype(F)
    return np.stack([col, col, col, np.ones_like(col)], -1)


# ---------------------------------------------------------------------------------------------- #
# Multi-program / multi-layer / temporal scenes of examples/basic/demo.py. A program renders
# each layer l of its texture with iLayer = l (shader.py:398-405); `X{t}x{l}` names the texture of t frames
# ago, layer l (texture.py:354-368); the program's own texture is RGBA8 unless the scene says otherwise.

def _opaque(rgb):
    return np.concatenate([rgb.astype(F), np.ones(rgb.shape[:-1] + (1,), F)], -1)


def frag_multishader_child(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/demo.py:74-79 (inline GLSL of MultiShader.child)"""
    z = np.zeros_like(f.stuv[..., 0])
    return _opaque(vec(z, (F(1) - f.stuv[..., 0]).astype(F), z))


def frag_multishader(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/demo.py:83-89 (inline GLSL of MultiShader.shader)"""
    z = np.zeros_like(f.stuv[..., 0])
    rgb = (vec(f.stuv[..., 0], z, z) + tex["child"].sample(f.astuv)[..., :3]).astype(F)
    return _opaque(rgb)


def _blur(image: Texture, stuv, radius: float, directions: int, steps: int):
    """multipass.frag:11-26 with its float loop counters evaluated in strict float32"""
    color = np.zeros(stuv.shape[:-1] + (4,), F)
    weights = F(0.0)
    direction, dstep, wstep = F(0.0), F(TAU/F(directions)), F(F(1.0)/F(steps))
    while direction < TAU:
        d = vec(np.cos(direction), np.sin(direction))
        walk = wstep
        while walk < F(1.0):
            offset = (((d*F(radius)).astype(F)*walk).astype(F)/F(2000)).astype(F)
            sample = image.sample((stuv + offset).astype(F))
            weight = F(F(1.0) - length(offset)/F(radius))
            color = (color + sample*weight).astype(F)
            weights = F(weights + weight)
            walk = F(walk + wstep)
        direction = F(direction + dstep)
    return (color/weights).astype(F)


def frag_multipass(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/multipass.frag:28-47"""
    if u.iLayer == 0:
        out = stexture(tex["background"], f.stuv)
    else:
        screen = tex["iScreen0x0"]
        out = screen.sample(f.astuv)
        left = out.copy(); left[..., 0] = F(1) - left[..., 0]
        out = np.where((f.gluv[..., 0] < 0)[..., None], left, _blur(screen, f.astuv, 5.0, 8, 8))
    out = out.astype(F).copy(); out[..., 3] = F(1)
    return out


def frag_motionblur(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/motionblur.frag:1-18; extra: iScreenTemporal"""
    cam = get_camera(u, f)
    if u.iLayer == 0:
        out = stexture(tex["background"], cam.stuv)
    else:
        T = int(u.extra["iScreenTemporal"])
        color = np.zeros(f.astuv.shape[:-1] + (4,), F)
        for i in range(T):
            factor = smoothstep(1.0, 0.0, F(i)/F(T))
            color = (color + tex[f"iScreen{i}x0"].sample(f.astuv)*factor).astype(F)
        out = ((F(2)*color)/F(T)).astype(F)
    out = out.astype(F).copy(); out[..., 3] = F(1)
    return out


def frag_dynamics(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/demo.py:120-125 (inline GLSL of Dynamics); extra: iShaderDynamics"""
    z = F(F(0.85) + F(0.1)*F(u.extra["iShaderDynamics"]))
    return stexture(tex["background"], zoom(f.stuv, np.full(f.stuv.shape[:-1], z, F), vec(F(0.5), F(0.5))))


def frag_audio(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/demo.py:150-154 (inline GLSL of Audio); extra: iAudioVolume"""
    v = np.full(f.stuv.shape[:-1], F(u.extra["iAudioVolume"]), F)
    return _opaque(vec(v, v, v))


LIFE_ALIVE = np.array([0, 0, 1, 1, 0, 0, 0, 0, 0], np.int32)    # life/simulation.glsl:7-11
LIFE_DEAD = np.array([0, 0, 0, 1, 0, 0, 0, 0, 0], np.int32)     # :14-18


def texel_fetch(tex: Texture, ix, iy):
    """texelFetch(sampler, ivec2, 0): no filtering, no wrapping; a fetch outside the image is undefined in
    GLSL 3.30 — fixed here as zeros (what robust buffer access returns)"""
    H, W, _ = tex.data.shape
    inside = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    t = tex.data[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)].astype(F)
    if tex.data.dtype == np.uint8:
        t = t/F(255)
    return np.where(inside[..., None], t, F(0))


def frag_life_simulation(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/life/simulation.glsl:20-50; extra: iLifePeriod, iLifeSize. Writes .r"""
    prev = tex["iLife1x0"]
    if (int(u.iFrame) % int(u.extra["iLifePeriod"])) != 0:
        r = prev.sample(f.astuv)[..., 0]
    else:
        size = np.asarray(u.extra["iLifeSize"], F)
        pixel = (f.astuv*size).astype(F).astype(np.int64)            # ivec2(): truncation
        near = np.zeros(pixel.shape[:-1], np.int32)
        current = np.zeros(pixel.shape[:-1], np.int32)
        for x in (-1, 0, 1):
            for y in (-1, 0, 1):
                cell = (texel_fetch(prev, pixel[..., 0] + x, pixel[..., 1] + y)[..., 0] > F(0.5)).astype(np.int32)
                if x == 0 and y == 0:
                    current = cell
                else:
                    near = near + cell
        r = np.where(current == 1, LIFE_ALIVE[near], LIFE_DEAD[near]).astype(F)
    z = np.zeros_like(r)
    return np.stack([r, z, z, np.ones_like(r)], -1).astype(F)


def frag_life_visuals(u: Uniforms, f: Frag, tex: dict):
    """examples/basic/shaders/life/visuals.glsl:11-38"""
    cam = get_camera(u, f)
    uv = cam.stuv
    exponent = F(1.3)
    area = F(F(1)/(exponent + F(1)))
    life = stexture(tex["iLife0x0"], uv)[..., 0]
    for k, base in ((1, 0.8), (2, 0.6), (3, 0.4), (4, 0.2)):
        life = (life + stexture(tex[f"iLife{k}x0"], uv)[..., 0]*gpow(F(base), exponent)).astype(F)
    life = (life/(F(5)*area)).astype(F)
    rgb = palette(life, *MAGMA)
    rgb = np.where(cam.out_of_bounds[..., None], np.asarray(MAGMA[0], F), rgb)
    return _opaque(rgb)


SCENES = dict(
    default=frag_default, shadertoy=frag_shadertoy, visualizer=frag_visualizer, bars=frag_bars,
    waveform=frag_waveform, mandelbrot=frag_mandelbrot, tetration=frag_tetration, raymarch=frag_raymarch,
    multishader_child=frag_multishader_child, multishader=frag_multishader, multipass=frag_multipass,
    motionblur=frag_motionblur, dynamics=frag_dynamics, audio=frag_audio,
    life_simulation=frag_life_simulation, life_visuals=frag_life_visuals,
)

# ---------------------------------------------------------------------------------------------- #
# Stores and the final pass

def to_unorm8(c):
    """Colour store to an 8-bit attachment"""
    with np.errstate(invalid="ignore"):
        q = np.rint(np.clip(np.nan_to_num(np.asarray(c, F), nan=0.0), 0, 1)*F(255))
    return q.astype(np.uint8)


def final_pass(screen_u8: np.ndarray, W: int, H: int, subsample: int) -> np.ndarray:
    """final.glsl:3-33 reading iScreen (RGBA8, LINEAR, CLAMP_TO_EDGE: scene.py:192-193), writing the
    W×H RGB8 target (scene.py:186-191). Returns float RGB before the 8-bit store"""
    u = Uniforms(iResolution=(W, H))
    f = varyings(u, W, H)
    scr = Texture(screen_u8, linear=True, repeat_x=False, repeat_y=False)
    if subsample == 1:
        return scr.sample(f.astuv)[..., :3]
    k = int(subsample)
    px = (F(1.0)/vec(F(W), F(H))).astype(F)
    corner = (f.astuv - px/F(2)).astype(F)
    origin = (corner + (px/F(k))/F(2)).astype(F)
    acc = np.zeros(f.astuv.shape[:-1] + (3,), F)
    for x in range(k):
        for y in range(k):
            off = ((px/F(k))*vec(F(x), F(y))).astype(F)
            acc = (acc + scr.sample((origin + off).astype(F))[..., :3]).astype(F)
    return (acc/F(k*k)).astype(F)


def render(scene: str, u: Uniforms, tex: dict, W: int, H: int, ssaa: float = 1.0, subsample: int = 2):
    """
    One exported frame: iScreen pass at the render resolution (scene.py:372-375), RGBA8 store,
    final.glsl downsample, RGB8 store. Arrays are bottom-row-first like fbo.read_into
    (exporting.py:165-174). Returns dict(screen_f32, screen_u8, final_f32, final_u8).
    """
    u.iResolution = (W, H)
    u.iSSAA = ssaa
    Wr, Hr = int(W*ssaa), int(H*ssaa)
    f = varyings(u, Wr, Hr)
    screen = SCENES[scene](u, f, tex)
    screen_u8 = to_unorm8(screen)
    fin = final_pass(screen_u8, W, H, subsample)
    return dict(screen_f32=screen, screen_u8=screen_u8, final_f32=fin, final_u8=to_unorm8(fin))

# ---------------------------------------------------------------------------------------------- #
# Synthetic assets (the shipped background is a network download, demo.py:29-37)

def synthetic_background(width: int = 1920, height: int = 1080, seed: int = 1) -> np.ndarray:
    """Deterministic RGB8 image, top row first (as PIL would load it): smooth colour field + grain"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    x /= width; y /= height
    img = np.stack([
        0.5 + 0.5*np.sin(6.0*x + 2.0*np.cos(5.0*y)),
        0.5 + 0.5*np.sin(4.0*y + 3.0*x*x + 1.0),
        0.5 + 0.5*np.cos(9.0*(x - 0.5)*(y - 0.5) + 2.0),
    ], -1)
    img += rng.uniform(-0.08, 0.08, img.shape)
    return (np.clip(img, 0, 1)*255).round().astype(np.uint8)
