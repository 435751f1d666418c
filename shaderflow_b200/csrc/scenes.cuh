// scenes.cuh — the reference's GLSL fragment shaders transliterated to CUDA device functions.
// Each `scene_*` is `void main()` of the cited file with fragColor returned; uniforms come from the
// kernel parameter block, varyings from `Frag` (vertex/default.glsl:1-17), samplers by slot
// (order reported by sfb_scene_info_get). Float loop counters are evaluated in strict IEEE float32.
#pragma once
#include "glsl.cuh"
#include "sampler.cuh"

namespace glsl {

using ::RenderParams;

struct Frag { vec2 agluv, gluv, astuv, stuv, stxy, glxy; };

// iAspectRatio macro (shaderflow.glsl:16)
SFB_HD float aspect_ratio(const sfb_uniforms& u) { return u.iResolution[0]/u.iResolution[1]; }

// The rasteriser: vertex/default.glsl:8-16 at the quad corners (shader.py:127-128), interpolated
// affinely (float64) to the centre of fragment (i, j), rounded once — same spec as the oracle.
SFB_HD Frag make_frag(const RenderParams& P, int i, int j) {
    const double tx = (double(i) + 0.5)*P.inv_Wr, ty = (double(j) + 0.5)*P.inv_Hr;
    const float W = P.u.iResolution[0], H = P.u.iResolution[1];
    const float a = aspect_ratio(P.u);
    auto lerp = [](double x0, double x1, double t) { return float(x0 + (x1 - x0)*t); };
    Frag f;
    f.agluv = mk2(lerp(-1.0, 1.0, tx), lerp(-1.0, 1.0, ty));
    f.gluv  = mk2(lerp(double(-1.0f*a), double(1.0f*a), tx), f.agluv.y);
    f.astuv = mk2(lerp(0.0, 1.0, tx), lerp(0.0, 1.0, ty));
    f.stuv  = mk2(lerp(double((-a + 1.0f)/2.0f), double((a + 1.0f)/2.0f), tx), f.astuv.y);
    f.stxy  = mk2(lerp(1.0, double(W*1.0f + 1.0f), tx), lerp(1.0, double(H*1.0f + 1.0f), ty));
    f.glxy  = mk2(lerp(double(1.0f - W/2.0f), double((W + 1.0f) - W/2.0f), tx),
                  lerp(double(1.0f - H/2.0f), double((H + 1.0f) - H/2.0f), ty));
    return f;
}

// ------------------------------------------------------------------------------------------------
// camera.glsl:55-155

struct Camera {
    vec2 gluv, agluv, stuv, astuv;
    vec3 origin, target;
    bool out_of_bounds;
};

SFB_HD vec3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }

SFB_HD Camera get_camera(const sfb_uniforms& u, const Frag& f) {
    const vec3 pos = ld3(u.iCameraPosition), right = ld3(u.iCameraRight), up = ld3(u.iCameraUpward);
    const vec3 fwd = ld3(u.iCameraForward), back = fwd*(-1.0f);
    const float asp = aspect_ratio(u);
    auto rectangle = [&](vec2 g, float size) { return size*(g.x*right + g.y*up); };                  // :55-57
    auto ray_origin = [&](vec3 p, vec2 g) {                                                           // :59-64
        return p + rectangle(g, u.iCameraZoom*u.iCameraIsometric) + (back*u.iCameraOrbital) + (back*u.iCameraDolly); };
    auto ray_target = [&](vec3 p, vec2 g) {                                                           // :66-71
        return p + rectangle(g, u.iCameraZoom) + (back*u.iCameraOrbital) + (fwd*u.iCameraFocalLength); };

    Camera c;
    if (u.iCameraProjection == 0) {
        c.origin = ray_origin(pos, f.gluv);
        c.target = ray_target(pos, f.gluv);
    } else if (u.iCameraProjection == 1) {                                                            // :101-110
        float sg = sign(f.agluv.x);
        vec2 g = f.gluv - sg*mk2(asp/2.0f, 0.0f);
        vec3 p = pos + (sg*u.iCameraSeparation)*right;
        c.origin = ray_origin(p, g);
        c.target = ray_target(p, g);
    } else {                                                                                          // :113-127
        float inclination = u.iCameraZoom*(PI*f.agluv.y/2.0f);
        float azimuth     = u.iCameraZoom*(PI*f.agluv.x/1.0f);
        vec3 t = rotate3d(fwd, right, -inclination);
        t = rotate3d(t, up, azimuth);
        c.origin = pos;
        c.target = pos + t;
    }
    // CameraRay2D (:73-91): plane through (0,0,1), normal (0,0,1)
    const vec3 pp = mk3(0.0f, 0.0f, 1.0f), pn = mk3(0.0f, 0.0f, 1.0f);
    float num = dot(pp - c.origin, pn);
    float den = dot(c.target - c.origin, pn);
    float t = num/den;
    c.out_of_bounds = (t < 0.0f) || (fabsf(f.gluv.x) > u.iWantAspect);
    vec3 hit = c.origin + (t*(c.target - c.origin));
    c.gluv  = xy(hit);
    c.agluv = c.gluv/mk2(asp, 1.0f);
    c.stuv  = gluv2stuv(c.gluv);
    c.astuv = gluv2stuv(c.agluv);
    return c;
}

// ------------------------------------------------------------------------------------------------
// resources/shaders/fragment/default.glsl:10-48

template <bool HW>
SFB_DEV vec4 scene_default(const RenderParams& P, const Frag& f) {
    Camera cam = get_camera(P.u, f);
    vec2 uv = cam.gluv;
    if (cam.out_of_bounds) return mk4(mk3(0.15f), 1.0f);
    vec3 rgb = mk3(0.0f);
    float angle = atan2_pos(uv.y, uv.x);
    vec3 color = 0.3f + hsv2rgb(angle + (2.0f*TAU*P.u.iTau) - (PI/4.0f), 1.0f, 1.0f);
    float circle = (1.333f*length(uv) - 1.0f);
    float width = 2.0f*fabsf(1.0f/(circle*circle))*1e-4f;
    if (circle < 0.0f) {
        rgb = rgb + mk3(0.18f);
    } else {
        const float grid = 8.0f;
        bool odd = mod(floorf(uv.x*grid/2.0f) + floorf(uv.y*grid/2.0f), 2.0f) > 0.5f;
        rgb = rgb + (odd ? mk3(0.22f) : mk3(0.20f));
    }
    rgb = rgb + (width*color);
    vec2 away = f.astuv*(1.0f - yx(f.astuv));
    float linear = 50.0f*(away.x*away.y);
    rgb = rgb*clamp(powf(linear, 0.1f), 0.0f, 1.0f);
    return mk4(rgb, 1.0f);
}

// examples/basic/shaders/shadertoy.frag:62-66
template <bool HW>
SFB_DEV vec4 scene_shadertoy(const RenderParams& P, const Frag& f) {
    float t = P.u.iTime;
    vec3 col = mk3(0.5f + 0.5f*cosf(t + f.stuv.x + 0.0f), 0.5f + 0.5f*cosf(t + f.stuv.y + 2.0f),
                   0.5f + 0.5f*cosf(t + f.stuv.x + 4.0f));
    return mk4(col, 1.0f);
}

// examples/basic/shaders/visualizer.frag:6-74
// extra[0].x = iAudioVolume, extra[1].x = iAudioSTD; samplers: background, iSpectrogram, iWaveform
template <bool HW>
SFB_DEV vec4 scene_visualizer(const RenderParams& P, const Frag& f) {
    const float iTime = P.u.iTime, iAudioVolume = P.u.extra[0][0], iAudioSTD = P.u.extra[1][0];
    const DevSampler& background = P.tex[0];
    Camera cam = get_camera(P.u, f);
    vec2 uv = cam.gluv;
    vec3 space = mk3(1.0f, 11.0f, 26.0f)/255.0f;
    if (cam.out_of_bounds) return mk4(space, 0.0f);

    vec2 background_uv = zoom(gluv2stuv(uv), 0.95f + 0.01f*sinf(iTime) - 0.02f*iAudioVolume - 0.03f, mk2(0.5f));
    background_uv = background_uv + 0.005f*mk2(cosf(iTime*3.25135f), sinf(iTime*1.153469f));
    vec4 fragColor = stexture<HW>(background, background_uv);
    {
        float intensity = 0.01f*clamp(powf(iAudioVolume, 2.5f), 0.0f, 0.3f);
        float quality = 10.0f, directions = 8.0f;
        vec4 color = fragColor;
        // Strict float32 counters: 9 angles × 10 walks
        for (float angle = 0.0f; angle < TAU; angle += TAU/directions) {
            vec2 dir = mk2(cosf(angle), sinf(angle));
            for (float walk = 1.0f/quality; walk <= 1.001f; walk += 1.0f/quality) {
                vec2 displacement = dir*walk*intensity;
                color = color + stexture<HW>(background, background_uv + displacement);
            }
        }
        fragColor = color/(quality*directions);
    }
    fragColor = fragColor*(1.0f + 5.0f*iAudioSTD*powf(clamp(length(f.agluv) - 0.3f, 0.0f, 1.0f), 6.0f));

    vec2 music_uv = rotate2d_mul(-PI/2.0f, uv);
    music_uv = music_uv*(1.0f - 0.4f*powf(fabsf(iAudioVolume), 0.5f));
    float radius = 0.17f;

    float circle = fabsf(atan1n(music_uv));
    vec4 s = texture<HW>(P.tex[1], mk2(0.0f, circle));
    vec2 freq = mk2(sqrtf(s.x/1000.0f), sqrtf(s.y/1000.0f));
    freq = freq*(0.05f + 3.0f*smoothstep(0.0f, 2.0f, circle));

    vec3 rgb = xyz(fragColor);
    if (length(music_uv) < radius) {
        rgb = rgb*0.5f;
    } else {
        float bar = (music_uv.y < 0.0f) ? freq.x : freq.y;
        float r = radius + 0.5f*bar;
        if (length(music_uv) < r) rgb = mix(rgb, mk3(1.0f), smoothstep(0.0f, 1.0f, 0.5f + bar));
        else                      rgb = rgb*powf((length(music_uv) - r)*0.5f, 0.05f);
    }
    rgb = mix(rgb, space, smoothstep(0.0f, 1.0f, length(uv)/20.0f));

    vec2 vig = f.astuv*(1.0f - yx(f.astuv));
    rgb = rgb*powf(vig.x*vig.y*20.0f, 0.1f + 0.15f*iAudioVolume);
    fragColor = mk4(rgb, 1.0f);

    vec4 w = texture<HW>(P.tex[2], mk2(f.astuv.x, 0.0f));
    vec2 wave = mk2(0.2f*w.x, 0.2f*w.y);
    if (1.0f - f.gluv.y < wave.x) fragColor = fragColor*0.8f;
    if (1.0f + f.gluv.y < wave.y) fragColor = fragColor*0.8f;
    return fragColor;
}

// examples/basic/shaders/bars.frag:5-22 — samplers: iSpectrogram
template <bool HW>
SFB_DEV vec4 scene_bars(const RenderParams& P, const Frag& f) {
    vec4 s = texture<HW>(P.tex[0], yx(f.astuv));
    vec2 intensity = mk2(sqrtf(s.x)/120.0f, sqrtf(s.y)/120.0f);
    vec3 rgb = mk3(0.0f);
    if (f.astuv.y < intensity.x) rgb = rgb + mk3(1.0f, 0.0f, 0.0f);
    if (f.astuv.y < intensity.y) rgb = rgb + mk3(0.0f, 1.0f, 0.0f);
    if (f.astuv.y < (intensity.y + intensity.x)/2.0f) rgb = rgb + mk3(0.0f, 0.0f, 1.0f);
    rgb = rgb + mk3(0.0f, 0.0f, 0.4f*(intensity.x + intensity.y)*(1.0f - f.astuv.y));
    return mk4(rgb, 1.0f);
}

// examples/basic/shaders/waveform.frag:5-19 — samplers: iWaveform
template <bool HW>
SFB_DEV vec4 scene_waveform(const RenderParams& P, const Frag& f) {
    vec4 w = texture<HW>(P.tex[0], mk2(f.astuv.x, 0.0f));
    vec4 c = mk4(mk3(0.2f), 1.0f);
    float ay = fabsf(f.gluv.y);
    if (ay < w.x) c.x = 1.0f;
    if (ay < w.y) c.y = 1.0f;
    if (ay < (w.x + w.y)/2.0f) c.z = 1.0f;
    return c;
}

// examples/fractals/shaders/mandelbrot.frag:10-31
// P.fast (every launch but SFB_RENDER_LITERAL): points of the main cardioid and of the period-2 disc never
// escape, so the loop would run to `quality` and leave iter == quality — the closed-form membership tests
// (with a 1e-3 safety margin towards the boundary, inside which the loop runs as written) skip those 500
// iterations for ~90 % of the interior. Same iter, same colour: tests hold the two variants to identical bytes.
template <bool HW>
SFB_DEV vec4 scene_mandelbrot(const RenderParams& P, const Frag& f) {
    Camera cam = get_camera(P.u, f);
    if (cam.out_of_bounds) return mk4(palette_magma(0.0f), 1.0f);
    vec2 z = cam.gluv - mk2(0.5f, 0.0f);
    vec2 c = z;
    int quality = int(1000.0f*P.u.iQuality);
    int iter = 0;
    bool inside = false;
    if (P.fast) {
        const float m = 1.0e-3f;
        const float px = (c.x - 0.25f)*(1.0f + m), py = c.y*(1.0f + m);          // scaled about the cusp (0.25, 0)
        const float q = px*px + py*py;
        const float bx = c.x + 1.0f;
        inside = (q*(q + px) < 0.25f*py*py) || (bx*bx + c.y*c.y < 0.0625f*(1.0f - m)*(1.0f - m));
    }
    if (inside) {
        iter = quality > 0 ? quality : 0;
    } else {
        for (; iter < quality; iter++) {
            // length(z) > 3.0 without the square root: sqrtf is correctly rounded and monotone, and the float
            // after 9 already has a root that rounds above 3, so the two tests agree for every dot(z, z)
            if (dot(z, z) > 9.0f) break;
            z = mk2(z.x*z.x - z.y*z.y, z.x*z.y + z.y*z.x) + c;
        }
    }
    float t = powf(1.0f - float(iter)/float(quality), 20.0f);
    return mk4(palette_magma(t), 1.0f);
}

// examples/fractals/shaders/tetration.frag:28-55
// The loop is pure issue-bound arithmetic (ncu: 94.7 % issue slots, no divergence: almost every fragment runs all
// 67 steps), ~98 instructions per step with libdevice powf + expf + logf + cosf + sinf. Two rewrites keep the
// mathematics and cut the instruction count:
//   * log(C.r) is loop-invariant (the GLSL recomputes it: same value); cos and sin of one angle come from one
//     sincosf (the same range reduction and polynomials as cosf / sinf) — every launch;
//   * P.fast (every launch but SFB_RENDER_LITERAL): pow(C.r, Z.x)·exp(−Z.y·C.t) = exp(Z.x·ln C.r − Z.y·C.t) is ONE
//     expf of the combined exponent, with ln C.r carried as a float-float pair so that Z.x·ln C.r keeps the
//     accuracy powf has internally; −Z.y·C.t is rounded to float32 first, exactly as the GLSL's argument of
//     exp is. One rounding instead of three (pow, exp, their product): within 1-2 ulp of the literal product.
//     The map is chaotic, so ulp-level differences move isolated pixels — measured on the oracle: 99.7 % of
//     channels within 1e-3 of the literal evaluation (a 1-ulp perturbation of pow alone gives 99.9 %); the
//     parity gate for this scene is 97 %.
template <bool HW>
SFB_DEV vec4 scene_tetration(const RenderParams& P, const Frag& f) {
    Camera cam = get_camera(P.u, f);
    float Cx = cam.gluv.x, Cy = cam.gluv.y;
    float Cr = sqrtf(Cx*Cx + Cy*Cy), Ct = atan2f(Cy, Cx);
    const float logCr = logf(Cr);
    float Zx = Cx, Zy = Cy, Zr = Cr;
    const int MAX_STEPS = 67;
    int it = 0;
    if (P.fast && Cr > 0.0f && Cr < 3.0e38f) {
        const double L = log(double(Cr));                  // once per fragment
        const float lh = float(L), ll = float(L - double(lh));
        for (it = 0; it < MAX_STEPS; it++) {
            const float e2 = -Zy*Ct;
            // exponent = Zx*(lh + ll) + e2: s is its float32 rounding (one FMA), lo what the rounding dropped —
            // e2 - s is exact whenever s and e2 are within a factor 2 (Sterbenz), else its rounding error is
            // below ulp(s)/2, i.e. no worse than the rounding of the literal product's exponent
            const float s = fmaf(Zx, lh, e2);
            const float lo = fmaf(Zx, ll, fmaf(Zx, lh, e2 - s));
            float r = expf(s);
            if (r < 3.0e38f) r = fmaf(r, lo, r);           // exp(s + lo) = exp(s)·(1 + lo + ...), |lo| < 1e-5; inf stays inf
            const float t = Zy*logCr + (Zx*Ct);
            float sn, cs;
            sincosf(t, &sn, &cs);
            Zr = r; Zx = r*cs; Zy = r*sn;
            if (Zr > 100.0f) break;
        }
    } else {
        for (it = 0; it < MAX_STEPS; it++) {
            float r = powf(Cr, Zx)*expf(-Zy*Ct);
            float t = Zy*logCr + (Zx*Ct);
            float sn, cs;
            sincosf(t, &sn, &cs);
            Zr = r; Zx = r*cs; Zy = r*sn;
            if (Zr > 100.0f) break;
        }
    }
    float k = float(it/MAX_STEPS);                 // integer division, tetration.frag:50
    float theta = atan2n(Zy, Zx);
    return mk4(hsv2rgb(theta, 1.0f, k), 1.0f);
}

// examples/basic/shaders/raymarch.frag:33-60
template <bool HW>
SFB_DEV vec4 scene_raymarch(const RenderParams& P, const Frag& f) {
    Camera cam = get_camera(P.u, f);
    vec3 origin = cam.origin;
    vec3 forward = normalize(cam.target - origin);
    float traveled = 0.0f, walk = 0.0f;
    int steps;
    for (steps = 0; steps < 100; steps++) {
        vec3 point = origin + (forward*traveled);
        float sdf = 2.0f*100.0f;
        for (int i = 2; i < 8; i++)
            sdf = fminf(sdf, sdBox(point, mk3(0.0f, 0.0f, float(i)), mk3(float(i - 1))));
        walk = sdf;
        traveled += walk;
        if (walk < 0.001f || walk > 100.0f) break;
    }
    return mk4(mk3(1.0f - sqrtf(float(steps))*0.1f), 1.0f);
}

// ------------------------------------------------------------------------------------------------
// Multi-program / multi-layer / temporal scenes of examples/basic/demo.py. The host renders
// each layer l of a program's texture with iLayer = l and rolls temporal textures afterwards
// (shader.py:398-405); samplers arrive in the order sfb_scene_info reports.

// examples/basic/demo.py:74-79 (inline GLSL of MultiShader.child)
template <bool HW>
SFB_DEV vec4 scene_multishader_child(const RenderParams& P, const Frag& f) {
    return mk4(0.0f, 1.0f - f.stuv.x, 0.0f, 1.0f);
}

// examples/basic/demo.py:83-89 (inline GLSL of MultiShader.shader) — samplers: child
template <bool HW>
SFB_DEV vec4 scene_multishader(const RenderParams& P, const Frag& f) {
    vec3 rgb = mk3(f.stuv.x, 0.0f, 0.0f);
    rgb = rgb + xyz(texture<HW>(P.tex[0], f.astuv));
    return mk4(rgb, 1.0f);
}

// multipass.frag:11-26 — float loop counters in strict IEEE float32, like the blur of visualizer.frag
template <bool HW>
SFB_DEV vec4 multipass_blur(const DevSampler& image, vec2 stuv, float radius, int directions, int steps) {
    vec4 color = mk4(0.0f);
    float weights = 0.0f;
    for (float direction = 0.0f; direction < TAU; direction += TAU/float(directions)) {
        for (float walk = 1.0f/float(steps); walk < 1.0f; walk += 1.0f/float(steps)) {
            const vec2 offset = ((mk2(cosf(direction), sinf(direction))*radius)*walk)/2000.0f;
            const vec4 sample = texture<HW>(image, stuv + offset);
            const float weight = 1.0f - length(offset)/radius;
            color = color + sample*weight;
            weights += weight;
        }
    }
    return color/weights;
}

// examples/basic/shaders/multipass.frag:28-47 — samplers: background, iScreen0x0
template <bool HW>
SFB_DEV vec4 scene_multipass(const RenderParams& P, const Frag& f) {
    vec4 c = mk4(0.0f);
    if (P.u.iLayer == 0) {
        c = stexture<HW>(P.tex[0], f.stuv);
    } else if (P.u.iLayer == 1) {
        c = texture<HW>(P.tex[1], f.astuv);
        if (f.gluv.x < 0.0f) c.x = 1.0f - c.x;
        else                 c = multipass_blur<HW>(P.tex[1], f.astuv, 5.0f, 8, 8);
    }
    c.w = 1.0f;
    return c;
}

// examples/basic/shaders/motionblur.frag:1-18 — extra[0].x = iScreenTemporal;
// samplers: background, iScreen0x0 ... iScreen{T-1}x0 (iScreenTexture(i, 0, uv), texture.py:363-367)
template <bool HW>
SFB_DEV vec4 scene_motionblur(const RenderParams& P, const Frag& f) {
    const Camera cam = get_camera(P.u, f);
    vec4 c = mk4(0.0f);
    if (P.u.iLayer == 0) {
        c = stexture<HW>(P.tex[0], cam.stuv);
    } else if (P.u.iLayer == 1) {
        const int T = min(int(P.u.extra[0][0]), SFB_MAX_SAMPLERS - 1);
        vec4 color = mk4(0.0f);
        for (int i = 0; i < T; i++) {
            const float factor = smoothstep(1.0f, 0.0f, float(i)/float(T));
            color = color + texture<HW>(P.tex[1 + i], f.astuv)*factor;
        }
        c = (2.0f*color)/float(T);
    }
    c.w = 1.0f;
    return c;
}

// examples/basic/demo.py:120-125 (inline GLSL of Dynamics) — extra[0].x = iShaderDynamics; samplers: background
template <bool HW>
SFB_DEV vec4 scene_dynamics(const RenderParams& P, const Frag& f) {
    const vec2 uv = zoom(f.stuv, 0.85f + 0.1f*P.u.extra[0][0], mk2(0.5f));
    return stexture<HW>(P.tex[0], uv);
}

// examples/basic/demo.py:150-154 (inline GLSL of Audio) — extra[0].x = iAudioVolume
template <bool HW>
SFB_DEV vec4 scene_audio(const RenderParams& P, const Frag& f) {
    return mk4(mk3(P.u.extra[0][0]), 1.0f);
}

// texelFetch(sampler, ivec2, 0).r: no filter, no wrap; outside the image GLSL 3.30 leaves the result
// undefined — zeros here, like the oracle (robust-access behaviour)
SFB_DEV float texel_fetch_r(const DevSampler& s, int ix, int iy) {
    if (ix < 0 || iy < 0 || ix >= s.w || iy >= s.h) return 0.0f;
    return texel_fetch(s, ix, iy).x;
}

// examples/basic/shaders/life/simulation.glsl:20-50 — extra[0].x = iLifePeriod, extra[1].xy = iLifeSize;
// samplers: iLife1x0 (the previous state). Only .r is meaningful (the target has one component)
template <bool HW>
SFB_DEV vec4 scene_life_simulation(const RenderParams& P, const Frag& f) {
    const int period = int(P.u.extra[0][0]);
    if (period != 0 && (P.u.iFrame % period) != 0)
        return mk4(texture<HW>(P.tex[0], f.astuv).x, 0.0f, 0.0f, 1.0f);
    const int px = int(f.astuv.x*P.u.extra[1][0]), py = int(f.astuv.y*P.u.extra[1][1]);
    int near = 0, current = 0;
    for (int x = -1; x <= 1; x++)
        for (int y = -1; y <= 1; y++) {
            const int cell = texel_fetch_r(P.tex[0], px + x, py + y) > 0.5f ? 1 : 0;
            if (x == 0 && y == 0) current = cell; else near += cell;
        }
    // alive[9] = {0,0,1,1,0,...}: two or three neighbours survive; dead[9] = {0,0,0,1,0,...}: three are born
    const int next = (current == 1) ? ((near == 2 || near == 3) ? 1 : 0) : ((near == 3) ? 1 : 0);
    return mk4(float(next), 0.0f, 0.0f, 1.0f);
}

// examples/basic/shaders/life/visuals.glsl:11-38 — samplers: iLife0x0 ... iLife4x0
template <bool HW>
SFB_DEV vec4 scene_life_visuals(const RenderParams& P, const Frag& f) {
    const Camera cam = get_camera(P.u, f);
    const vec3 C1 = mk3(0.01060815f, 0.01808215f, 0.10018654f), C2 = mk3(0.38092887f, 0.12061482f, 0.32506528f);
    const vec3 C3 = mk3(0.79650140f, 0.10506637f, 0.31063031f), C4 = mk3(0.95922872f, 0.53307513f, 0.37488950f);
    if (cam.out_of_bounds) return mk4(C1, 1.0f);
    const vec2 uv = cam.stuv;
    const float exponent = 1.3f;
    const float area = 1.0f/(exponent + 1.0f);
    float life = 0.0f;
    life += stexture<HW>(P.tex[0], uv).x;
    life += stexture<HW>(P.tex[1], uv).x*powf(0.8f, exponent);
    life += stexture<HW>(P.tex[2], uv).x*powf(0.6f, exponent);
    life += stexture<HW>(P.tex[3], uv).x*powf(0.4f, exponent);
    life += stexture<HW>(P.tex[4], uv).x*powf(0.2f, exponent);
    life /= (5.0f*area);
    return mk4(palette(life, C1, C2, C3, C4), 1.0f);
}

// texelFetch(sampler, ivec2, 0): all four components (zeros outside, like texel_fetch_r)
SFB_DEV vec4 texel_fetch_raw(const DevSampler& s, int ix, int iy) {
    if (ix < 0 || iy < 0 || ix >= s.w || iy >= s.h) return mk4(0.0f);
    return texel_fetch(s, ix, iy);
}

// examples/shaders/piano.frag of this repository (the reference ships ShaderPiano, piano/module.py, without a
// fragment). extra: iPianoDynamic (vec2), iPianoExtra, iPianoHeight, iPianoBlackRatio, iPianoRollTime,
// iPianoLimit; samplers: iPianoKeys, iPianoChan, iPianoRoll
SFB_DEV vec3 piano_channel_color(float channel) { return hsv2rgb(0.6f + 0.9f*channel, 0.75f, 1.0f); }

template <bool HW>
SFB_DEV vec4 scene_piano(const RenderParams& P, const Frag& f) {
    const float dyn_lo = P.u.extra[0][0], dyn_hi = P.u.extra[0][1], extra = P.u.extra[1][0];
    const float height = P.u.extra[2][0], black_ratio = P.u.extra[3][0], roll_time = P.u.extra[4][0];
    const int limit = int(P.u.extra[5][0]);
    vec3 rgb = mk3(0.06f);
    const float lo = dyn_lo - extra, hi = dyn_hi + extra + 1.0f;
    const float keyf = mix(lo, hi, f.astuv.x);
    const float keyi = floorf(keyf);
    const float inkey = keyf - keyi;
    if ((keyi < 0.0f) || (keyi > 127.0f)) return mk4(rgb, 1.0f);
    const int key = int(keyi);
    const int k12 = key - 12*(key/12);
    const bool black = (k12 == 1) || (k12 == 3) || (k12 == 6) || (k12 == 8) || (k12 == 10);
    if (f.astuv.y < height) {
        const float press = clamp(texel_fetch_raw(P.tex[0], key, 0).x/128.0f, 0.0f, 1.0f);
        const float chan = texel_fetch_raw(P.tex[1], key, 0).x;
        vec3 base = black ? mk3(0.12f) : mk3(0.92f);
        if (black && (f.astuv.y < height*(1.0f - black_ratio))) base = mk3(0.92f);
        rgb = base;
        if (chan >= 0.0f) rgb = mix(base, piano_channel_color(chan), press);
        if (inkey < 0.06f) rgb = rgb*0.55f;
    } else {
        const float when = P.u.iTime + roll_time*(f.astuv.y - height)/(1.0f - height);
        rgb = black ? mk3(0.09f) : mk3(0.12f);
        for (int i = 0; i < limit; i++) {
            const vec4 note = texel_fetch_raw(P.tex[2], i, key);
            if (note.x == 0.0f && note.y == 0.0f && note.z == 0.0f && note.w == 0.0f) break;
            if ((note.x <= when) && (when <= note.y)) {
                rgb = piano_channel_color(note.z)*(0.35f + 0.65f*note.w/128.0f);
                if ((inkey < 0.08f) || (inkey > 0.92f)) rgb = rgb*0.5f;
            }
        }
    }
    return mk4(rgb, 1.0f);
}

// ------------------------------------------------------------------------------------------------
// visualizer.frag, per-fragment fast path (used where neither visualizer_rows.cu nor visualizer_tiled.cuh
// applies: float32 debug outputs, targets that are not RGBA8, CTAs whose window does not fit). Same
// mathematics as scene_visualizer (the literal transliteration above stays as the parity anchor and serves
// SFB_FILTER_HARDWARE / unusual texture formats); what
// changes is how the 91 bilinear taps of the blur loop (visualizer.frag:19-33) are evaluated:
//   * the (angle, walk) pairs of the two strict-float32 loops are a 90-entry constant table
//     (dir*walk), built on the host with the same float arithmetic (render.cu: build_blur_table);
//   * stexture() on the RGBA8 background is specialised: no integer modulo unless the 2x2 footprint
//     touches the texture edge, floor via the 1.5*2^23 trick, bytes widened with PRMT + FADD instead of
//     I2F, the 1/255 of the unorm decode applied once to the tap sum;
//   * consecutive taps along a walk usually land in the same texel quad: the unpacked quad is kept in
//     registers and refetched only when (i0, j0) changes.
// Differences to the literal path are float32 re-association only (~1e-7 relative).

struct BlurTable { float2 tap[92]; };   // [0,90): dir*walk; [90] = (0,0), the undisplaced first tap
__constant__ BlurTable c_blur;

struct QuadSampler {
    const uchar4* texels; int w, h; int rx, ry;
    float fw, fh, hw;
    int qi, qj;
    vec3 t00, t10, t01, t11;

    SFB_DEV void init(const DevSampler& s) {
        texels = reinterpret_cast<const uchar4*>(s.lin); w = s.w; h = s.h; rx = s.rx; ry = s.ry;
        fw = float(s.w); fh = float(s.h); hw = float(s.h)/float(s.w);
        qi = qj = 0x7fffffff;
    }
    static SFB_DEV vec3 widen(uchar4 c) {
        const unsigned int word = (unsigned int)c.x | ((unsigned int)c.y << 8) | ((unsigned int)c.z << 16) | ((unsigned int)c.w << 24);
        // byte k into the mantissa of 2^23: float = 8388608 + byte
        return mk3(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540)) - 8388608.0f,
                   __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7541)) - 8388608.0f,
                   __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7542)) - 8388608.0f);
    }
    SFB_DEV void fetch(int i0, int j0) {
        int i1 = i0 + 1, j1 = j0 + 1;
        if ((unsigned int)i0 >= (unsigned int)(w - 1) || (unsigned int)j0 >= (unsigned int)(h - 1)) {
            i0 = wrap_index(i0, w, rx); i1 = wrap_index(i1, w, rx);
            j0 = wrap_index(j0, h, ry); j1 = wrap_index(j1, h, ry);
        }
        const uchar4* r0 = texels + size_t(j0)*size_t(w);
        const uchar4* r1 = texels + size_t(j1)*size_t(w);
        t00 = widen(__ldg(r0 + i0)); t10 = widen(__ldg(r0 + i1));
        t01 = widen(__ldg(r1 + i0)); t11 = widen(__ldg(r1 + i1));
    }
    // stexture(image, st) in byte units (0..255); shaderflow.glsl:165-169,202-204 + GL bilinear
    SFB_DEV vec3 stexture255(vec2 st) {
        const float ux = ((st.x*2.0f - 1.0f)*hw + 1.0f)/2.0f;
        const float uy = ((st.y*2.0f - 1.0f) + 1.0f)/2.0f;
        const float ub = ux*fw - 0.5f, vb = uy*fh - 0.5f;
        // floor for |x| < 2^22: round to nearest through the magic constant, step down if it went up
        const float MAGIC = 12582912.0f;
        float mx = ub + MAGIC, my = vb + MAGIC;
        float fx = mx - MAGIC, fy = my - MAGIC;
        if (fx > ub) { fx -= 1.0f; mx -= 1.0f; }
        if (fy > vb) { fy -= 1.0f; my -= 1.0f; }
        const int i0 = __float_as_int(mx) - 0x4B400000, j0 = __float_as_int(my) - 0x4B400000;
        const float a = ub - fx, b = vb - fy;
        if (i0 != qi || j0 != qj) { fetch(i0, j0); qi = i0; qj = j0; }
        const vec3 top = t00*(1.0f - a) + t10*a;
        const vec3 bot = t01*(1.0f - a) + t11*a;
        return top*(1.0f - b) + bot*b;
    }
};

SFB_DEV vec4 scene_visualizer_fast(const RenderParams& P, const Frag& f) {
    const float iTime = P.u.iTime, iAudioVolume = P.u.extra[0][0], iAudioSTD = P.u.extra[1][0];
    Camera cam = get_camera(P.u, f);
    vec2 uv = cam.gluv;
    vec3 space = mk3(1.0f, 11.0f, 26.0f)/255.0f;
    if (cam.out_of_bounds) return mk4(space, 0.0f);

    vec2 background_uv = zoom(gluv2stuv(uv), 0.95f + 0.01f*sinf(iTime) - 0.02f*iAudioVolume - 0.03f, mk2(0.5f));
    background_uv = background_uv + 0.005f*mk2(cosf(iTime*3.25135f), sinf(iTime*1.153469f));
    QuadSampler bg; bg.init(P.tex[0]);
    vec3 sum = bg.stexture255(background_uv);
    {
        const float intensity = 0.01f*clamp(powf(iAudioVolume, 2.5f), 0.0f, 0.3f);
        #pragma unroll 10
        for (int t = 0; t < 90; t++) {
            const float2 d = c_blur.tap[t];
            sum = sum + bg.stexture255(background_uv + mk2(d.x, d.y)*intensity);
        }
    }
    // (first tap + 90 taps) / (quality*directions), unorm decode folded in
    vec3 rgb = sum*((1.0f/255.0f)/(10.0f*8.0f));
    rgb = rgb*(1.0f + 5.0f*iAudioSTD*powf(clamp(length(f.agluv) - 0.3f, 0.0f, 1.0f), 6.0f));

    // rotate2d(-PI/2): cos = -4.371139e-08, sin = -1 in float32
    const float c = -4.37113883e-08f, sn = -1.0f;
    vec2 music_uv = mk2(c*uv.x + sn*uv.y, (-sn)*uv.x + c*uv.y);
    music_uv = music_uv*(1.0f - 0.4f*powf(fabsf(iAudioVolume), 0.5f));
    const float radius = 0.17f;
    const float circle = fabsf(atan1n(music_uv));
    vec4 s = texture<false>(P.tex[1], mk2(0.0f, circle));
    vec2 freq = mk2(sqrtf(s.x/1000.0f), sqrtf(s.y/1000.0f));
    freq = freq*(0.05f + 3.0f*smoothstep(0.0f, 2.0f, circle));
    const float lm = length(music_uv);
    if (lm < radius) {
        rgb = rgb*0.5f;
    } else {
        const float bar = (music_uv.y < 0.0f) ? freq.x : freq.y;
        const float r = radius + 0.5f*bar;
        if (lm < r) rgb = mix(rgb, mk3(1.0f), smoothstep(0.0f, 1.0f, 0.5f + bar));
        else        rgb = rgb*powf((lm - r)*0.5f, 0.05f);
    }
    rgb = mix(rgb, space, smoothstep(0.0f, 1.0f, length(uv)/20.0f));
    vec2 vig = f.astuv*(1.0f - yx(f.astuv));
    rgb = rgb*powf(vig.x*vig.y*20.0f, 0.1f + 0.15f*iAudioVolume);
    vec4 fragColor = mk4(rgb, 1.0f);
    vec4 w = texture<false>(P.tex[2], mk2(f.astuv.x, 0.0f));
    if (1.0f - f.gluv.y < 0.2f*w.x) fragColor = fragColor*0.8f;
    if (1.0f + f.gluv.y < 0.2f*w.y) fragColor = fragColor*0.8f;
    return fragColor;
}

template <int SCENE, bool HW>
SFB_DEV vec4 shade(const RenderParams& P, const Frag& f) {
    if constexpr (SCENE == SFB_SCENE_DEFAULT)    return scene_default<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_SHADERTOY)  return scene_shadertoy<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_VISUALIZER) {
        if (!HW && P.fast) return scene_visualizer_fast(P, f);
        return scene_visualizer<HW>(P, f);
    }
    if constexpr (SCENE == SFB_SCENE_BARS)       return scene_bars<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_WAVEFORM)   return scene_waveform<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_MANDELBROT) return scene_mandelbrot<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_TETRATION)  return scene_tetration<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_RAYMARCH)   return scene_raymarch<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_MULTISHADER_CHILD) return scene_multishader_child<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_MULTISHADER)       return scene_multishader<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_MULTIPASS)         return scene_multipass<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_MOTIONBLUR)        return scene_motionblur<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_DYNAMICS)          return scene_dynamics<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_AUDIO)             return scene_audio<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_LIFE_SIMULATION)   return scene_life_simulation<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_LIFE_VISUALS)      return scene_life_visuals<HW>(P, f);
    if constexpr (SCENE == SFB_SCENE_PIANO)             return scene_piano<HW>(P, f);
    return mk4(0.0f);
}

} // namespace glsl
