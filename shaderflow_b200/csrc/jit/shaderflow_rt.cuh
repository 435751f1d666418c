// shaderflow_rt.cuh — what a ShaderFlow fragment shader finds around its `main()`, for run-time compiled programs:
//   ShaderBase   the built-in uniforms (scene.py:687-703, camera.py:146-201 → sfb_uniforms), the varyings of
//                vertex/default.glsl:1-17 at this fragment (the rasteriser rule of scenes.cuh make_frag), fragColor,
//                and ShaderFlow's GLSL std-lib API (shaderflow/resources/shaders/include/shaderflow.glsl and
//                camera.glsl: same names, arguments and arithmetic, written against g::vec) so user shaders that call
//                it compile unchanged. The translator emits `struct Shader : g::ShaderBase { ... }`: a member a user
//                shader defines itself hides the one here, as a later definition would in the GLSL header.
//   kernels      sfb_jit_screen (one thread per fragment, any target format — screen_kernel of render_kernels.cuh)
//                and sfb_jit_frame (fused ssaa×ssaa + final box, rgb24 / rgba8 — frame_kernel), same stores, same
//                8-bit rules. Included after the emitted `Shader`.
// NVRTC only. Included after sfb200.h, render_params.h and glsl_rt.cuh.
#pragma once

namespace g {

constexpr float PI = 3.1415926535897932f, TAU = 6.2831853071795864f;
constexpr float SQRT2 = 1.4142135623730951f, SQRT3 = 1.7320508075688772f, SQRT5 = 2.2360679774997898f;
constexpr int CameraModeFreeCamera = 0, CameraMode2D = 1, CameraModeSpherical = 2;
constexpr int CameraProjectionPerspective = 0, CameraProjectionStereoscopic = 1, CameraProjectionEquirectangular = 2;

struct Camera {
    int mode, projection;
    vec3 position, up, down, left, right, forward, backward, zenith;
    vec3 origin, target;
    float orbital, dolly;
    vec3 plane_point, plane_normal;
    vec2 gluv, agluv, stuv, astuv, glxy, stxy;
    bool out_of_bounds;
    float separation, focal_length, isometric, zoom;
};

// --- std-lib entries that read no uniform (shaderflow.glsl line numbers cited per group)

G_DEV float proportion(float a, float b, float c) { return (b*c)/a; }                                              // :24
G_DEV float lerp(float ax, float ay, float bx, float by, float x) { return ay + (x - ax)*(by - ay)/(bx - ax); }    // :29
G_DEV float smoothlerp(float a, float b, float difference) {                                                       // :36-40
    const float t = clamp((a - b)/difference + 0.5f, 0.0f, 1.0f);
    const float offset = difference*t*(1.0f - t)/2.0f;
    return mix(a, b, t) - offset;
}
G_DEV float smin(float a, float b, float k) { return smoothlerp(a, b, k); }                                        // :43-46
G_DEV float smax(float a, float b, float k) { return smoothlerp(a, b, -k); }
G_DEV float smin(float a, float b) { return smoothlerp(a, b, 1.0f); }
G_DEV float smax(float a, float b) { return smoothlerp(a, b, -1.0f); }
G_DEV float smoothmix(float a, float b, float x0, float x1, float x) { return mix(a, b, smoothstep(x0, x1, x)); }  // :50
G_DEV float smix(float a, float b, float x0, float x1, float x) { return smoothmix(a, b, x0, x1, x); }             // :55
G_DEV float triangle_wave(float x, float period) { return 2.0f*abs(mod(2.0f*x/period - 0.5f, 2.0f) - 1.0f) - 1.0f; }   // :62
template <int N> G_DEV float angle(const vec<float, N>& a, const vec<float, N>& b) { return acos(dot(a, b)/(length(a)*length(b))); }   // :70-72
G_DEV mat2 rotate2d(float a) { return mat2(cos(a), -sin(a), sin(a), cos(a)); }                                     // :75-77
G_DEV mat2 rotate2deg(float a) { return rotate2d(radians(a)); }
G_DEV vec3 rotate3d(vec3 v, vec3 axis, float a) { return mix(dot(axis, v)*axis, v, cos(a)) + cross(axis, v)*sin(a); }  // :82-84
G_DEV vec3 rotate3deg(vec3 v, vec3 axis, float a) { return rotate3d(v, axis, radians(a)); }
G_DEV vec2 stuv2gluv(vec2 s) { return (s*2) - 1; }                                                                 // :91-96
G_DEV vec2 s2g(vec2 s) { return stuv2gluv(s); }
G_DEV vec2 gluv2stuv(vec2 p) { return (p + 1)/2; }
G_DEV vec2 g2s(vec2 p) { return gluv2stuv(p); }
G_DEV vec2 stuv2stxy(vec2 s, vec2 resolution) { return resolution*s; }                                             // :103
G_DEV vec2 stxy2stuv(vec2 s, vec2 resolution) { return s/resolution; }                                             // :107
G_DEV vec2 agluv_mirrored_repeat(vec2 p) { return vec2(triangle_wave(p.x, 4.0f), triangle_wave(p.y, 4.0f)); }      // :119
G_DEV bool astuv_oob(vec2 p) { return (p.x < 0) || (p.x > 1) || (p.y < 0) || (p.y > 1); }                          // :135
G_DEV bool agluv_oob(vec2 p) { return (p.x < -1) || (p.x > 1) || (p.y < -1) || (p.y > 1); }                        // :141
G_DEV vec2 polar2rect(float radius, float a) { return radius*vec2(cos(a), sin(a)); }                               // :149
G_DEV vec3 sphere2rect(float radius, float theta, float phi) {                                                     // :154
    return vec3(radius*sin(theta)*cos(phi), radius*sin(theta)*sin(phi), radius*cos(theta)); }
G_DEV vec4 gtexture(sampler2D image, vec2 p) {                                                                     // :165-169
    const vec2 resolution = textureSize(image, 0);
    const vec2 scale = vec2(resolution.y/resolution.x, 1);
    return texture(image, gluv2stuv(p*scale));
}
G_DEV vec4 stexture(sampler2D image, vec2 s) { return gtexture(image, stuv2gluv(s)); }                             // :202
G_DEV vec4 astexture(sampler2D image, vec2 s) { return texture(image, s); }                                        // :206
G_DEV vec3 palette(float t, vec3 A, vec3 B, vec3 C, vec3 D) {                                                      // :212-220
    if (t < 0.25f) return mix(A, B, t*4);
    else if (t < 0.5f) return mix(B, C, (t - 0.25f)*4);
    return mix(C, D, (t - 0.5f)*4);
}
G_DEV vec3 palette_magma(float t) {                                                                                // :222-226
    return palette(t, vec3(0.01060815f, 0.01808215f, 0.10018654f), vec3(0.38092887f, 0.12061482f, 0.32506528f),
                      vec3(0.79650140f, 0.10506637f, 0.31063031f), vec3(0.95922872f, 0.53307513f, 0.37488950f));
}
G_DEV bool isBlackKey(int index) { const int key = index % 12; return key == 1 || key == 3 || key == 6 || key == 8 || key == 10; }   // :231-245
G_DEV bool isBlackKey(float key) { return isBlackKey(int(key)); }
G_DEV bool isWhiteKey(int index) { return !isBlackKey(index); }
G_DEV bool isWhiteKey(float key) { return isWhiteKey(int(key)); }
G_DEV float _sdLine(vec3 origin, vec3 A, vec3 B, bool segment) {                                                   // :255-261
    const vec3 direction = B - A, shortest = origin - A;
    float t = dot(shortest, direction)/dot(direction, direction);
    if (segment) t = clamp(t, 0, 1);
    return length(shortest - direction*t);
}
G_DEV float sdLine(vec2 o, vec2 p1, vec2 p2) { return _sdLine(vec3(o, 0), vec3(p1, 0), vec3(p2, 0), false); }      // :263-271
G_DEV float sdLine(vec3 o, vec3 p1, vec3 p2) { return _sdLine(o, p1, p2, false); }
G_DEV float sdLineSegment(vec3 o, vec3 p1, vec3 p2) { return _sdLine(o, p1, p2, true); }
G_DEV float sdLineSegment(vec2 o, vec2 p1, vec2 p2) { return _sdLine(vec3(o, 0), vec3(p1, 0), vec3(p2, 0), true); }
G_DEV float sdSphere(vec3 origin, vec3 position, float radius) { return length(position - origin) - radius; }      // :275
G_DEV float sdPlane(vec3 origin, vec3 point, vec3 normal) { return dot(origin - point, normalize(normal)); }       // :280
G_DEV float sdBox(vec3 origin, vec3 point, vec3 size) {                                                            // :285-288
    const vec3 d = abs(origin - point) - size/2;
    return min(max(d.x, max(d.y, d.z)), 0) + length(max(d, 0));
}
G_DEV float sdOctahedron(vec3 origin, vec3 point, float size) { const vec3 p = abs(origin - point); return SQRT3*(p.x + p.y + p.z - size); }   // :291
G_DEV float sdUnion(float a, float b) { return min(a, b); }                                                        // :299-332
G_DEV float sdSmoothUnion(float a, float b, float width) { const float k = clamp(0.5f + 0.5f*(b - a)/width, 0, 1); return mix(b, a, k) - width*k*(1 - k); }
G_DEV float sdSubtraction(float a, float b) { return max(b, -a); }
G_DEV float sdSmoothSubtraction(float a, float b, float width) { const float k = clamp(0.5f - 0.5f*(b + a)/width, 0, 1); return mix(b, -a, k) + width*k*(1 - k); }
G_DEV float sdIntersection(float a, float b) { return max(a, b); }
G_DEV float sdSmoothIntersection(float a, float b, float width) { const float k = clamp(0.5f - 0.5f*(b - a)/width, 0, 1); return mix(b, a, k) + width*k*(1 - k); }
G_DEV vec4 blend(vec4 a, vec4 b) { return mix(a, b, b.a); }                                                        // :343
G_DEV vec4 alpha_composite(vec4 a, vec4 b) { return a*(1 - b.a) + (b*b.a); }                                       // :349
template <int N> G_DEV vec<float, N> saturate(const vec<float, N>& color, float amount) { return clamp(color*amount, 0, 1); }   // :354-356
G_DEV vec2 zoom(vec2 uv, float z, vec2 anchor) { return (uv - anchor)*(z*z) + anchor; }                            // :361-367
G_DEV vec2 zoom(vec2 uv, float z) { return uv*(z*z); }
G_DEV float atan_normalized(float x) { return 2*atan(x)/PI; }                                                      // :370-400
G_DEV float atan1(vec2 p) { return atan(p.y, p.x); }
G_DEV float atan1n(vec2 p) { return atan(p.y, p.x)/PI; }
G_DEV float atan2(float y, float x) { if (y < 0) return TAU - atan(-y, x); return atan(y, x); }
G_DEV float atan2(vec2 p) { return atan2(p.y, p.x); }
G_DEV float atan2n(float y, float x) { return atan2(y, x)/TAU; }
G_DEV float atan2n(vec2 p) { return atan2n(p.y, p.x); }
G_DEV vec3 hsv2rgb(vec3 hsv) {                                                                                     // :406-425
    float h = hsv.x; const float s = hsv.y, v = hsv.z;
    h = mod(h, TAU);
    const float c = v*s, x = c*(1 - abs(mod(h/(PI/3), 2) - 1)), m = v - c;
    vec3 rgb = vec3(0.5f);
    switch (int(floor(6*(h/(2*PI))))) {
        case 0: rgb = vec3(c, x, 0); break;
        case 1: rgb = vec3(x, c, 0); break;
        case 2: rgb = vec3(0, c, x); break;
        case 3: rgb = vec3(0, x, c); break;
        case 4: rgb = vec3(x, 0, c); break;
        case 5: rgb = vec3(c, 0, x); break;
        default: rgb = vec3(0);
    }
    return rgb + vec3(m);
}
G_DEV vec3 hsv2rgb(float h, float s, float v) { return hsv2rgb(vec3(h, s, v)); }
G_DEV vec4 hsv2rgb(vec4 hsv) { return vec4(hsv2rgb(swz<0, 1, 2>(hsv)), hsv.a); }
G_DEV vec3 rgb2hsv(vec3 rgb) {                                                                                     // :432-450
    const float cmax = max(rgb.r, max(rgb.g, rgb.b)), cmin = min(rgb.r, min(rgb.g, rgb.b)), delta = cmax - cmin;
    float h = 0;
    if (delta == 0) h = 0;
    else if (cmax == rgb.r) h = mod((rgb.g - rgb.b)/delta, 6);
    else if (cmax == rgb.g) h = (rgb.b - rgb.r)/delta + 2;
    else h = (rgb.r - rgb.g)/delta + 4;
    h *= PI/3;
    const float s = (cmax == 0) ? 0.0f : delta/cmax;
    return vec3(h, s, cmax);
}
G_DEV vec3 rgb2hsv(float r, float g_, float b) { return rgb2hsv(vec3(r, g_, b)); }
G_DEV vec4 rgb2hsv(vec4 rgb) { return vec4(rgb2hsv(swz<0, 1, 2>(rgb)), rgb.a); }
G_DEV float noise21(vec2 p) { return fract(sin(dot(p, vec2(18.4835183f, 59.583596f)))*39758.381532f); }           // :455-466
G_DEV vec2 noise22(vec2 p) { const float x = noise21(p); return vec2(x, noise21(p + x)); }
G_DEV float noise11(float f) { return fract(sin(f)*39758.381532f); }

// --- everything that reads uniforms or varyings

struct ShaderBase {
    // scene.py:687-703
    float iTime, iTau, iDuration, iDeltatime, iFrametime, iCycle, iWantAspect, iQuality, iSSAA, iFramerate, iAspectRatio, iWidth, iHeight;
    vec2 iResolution, iMouse;
    int iFrame, iLayer;
    bool iRealtime, iRendering, iMouseInside, iMouse1, iMouse2;
    // camera.py:146-201
    int iCameraMode, iCameraProjection;
    vec3 iCameraPosition, iCameraRight, iCameraUpward, iCameraForward, iCameraZenith;
    float iCameraZoom, iCameraIsometric, iCameraFocalLength, iCameraOrbital, iCameraDolly, iCameraSeparation;
    // vertex/default.glsl:1-17
    vec2 fragCoord, stxy, glxy, stuv, astuv, gluv, agluv;
    int instance;
    vec4 gl_FragCoord;
    vec4 fragColor;
    bool sfb_discarded;
    float sfb_dscale;                 // fragments between this lane and its quad neighbours, inverted (jit_kernels.cuh)
    const RenderParams* sfb_params;

    G_DEV static float lerp64(double x0, double x1, double t) { return float(x0 + (x1 - x0)*t); }

    // uniforms of the block + the rasteriser: every `out` of the vertex shader is affine over the quad, evaluated in
    // float64 at the centre of fragment (i, j) and rounded once (the rule of scenes.cuh make_frag and of the oracle)
    G_DEV ShaderBase(const RenderParams& P, int i, int j) {
        const sfb_uniforms& u = P.u;
        sfb_params = &P;
        iTime = u.iTime; iTau = u.iTau; iDuration = u.iDuration; iWantAspect = u.iWantAspect; iQuality = u.iQuality;
        iSSAA = u.iSSAA; iFramerate = u.iFramerate;
        iFrametime = 1.0f/iFramerate; iDeltatime = iFrametime;              // shaderflow.glsl:13-14 (the macro shadows the uniform)
        iCycle = (2*PI)*iTau;                                               // :15
        iResolution = vec2(u.iResolution[0], u.iResolution[1]);
        iAspectRatio = iResolution.x/iResolution.y; iWidth = iResolution.x; iHeight = iResolution.y;   // :16-18
        iMouse = vec2(u.iMouse[0], u.iMouse[1]);
        iFrame = u.iFrame; iLayer = u.iLayer;
        iRealtime = u.iRealtime != 0; iRendering = !iRealtime; iMouseInside = u.iMouseInside != 0; iMouse1 = u.iMouse1 != 0; iMouse2 = u.iMouse2 != 0;
        iCameraMode = u.iCameraMode; iCameraProjection = u.iCameraProjection;
        iCameraPosition = vec3(u.iCameraPosition[0], u.iCameraPosition[1], u.iCameraPosition[2]);
        iCameraRight = vec3(u.iCameraRight[0], u.iCameraRight[1], u.iCameraRight[2]);
        iCameraUpward = vec3(u.iCameraUpward[0], u.iCameraUpward[1], u.iCameraUpward[2]);
        iCameraForward = vec3(u.iCameraForward[0], u.iCameraForward[1], u.iCameraForward[2]);
        iCameraZenith = vec3(u.iCameraZenith[0], u.iCameraZenith[1], u.iCameraZenith[2]);
        iCameraZoom = u.iCameraZoom; iCameraIsometric = u.iCameraIsometric; iCameraFocalLength = u.iCameraFocalLength;
        iCameraOrbital = u.iCameraOrbital; iCameraDolly = u.iCameraDolly; iCameraSeparation = u.iCameraSeparation;

        const double tx = (double(i) + 0.5)*P.inv_Wr, ty = (double(j) + 0.5)*P.inv_Hr;
        const float W = iResolution.x, H = iResolution.y, a = iAspectRatio;
        agluv = vec2(lerp64(-1.0, 1.0, tx), lerp64(-1.0, 1.0, ty));
        gluv  = vec2(lerp64(double(-1.0f*a), double(1.0f*a), tx), agluv.y);
        astuv = vec2(lerp64(0.0, 1.0, tx), lerp64(0.0, 1.0, ty));
        stuv  = vec2(lerp64(double((-a + 1.0f)/2.0f), double((a + 1.0f)/2.0f), tx), astuv.y);
        stxy  = vec2(lerp64(1.0, double(W*1.0f + 1.0f), tx), lerp64(1.0, double(H*1.0f + 1.0f), ty));
        glxy  = vec2(lerp64(double(1.0f - W/2.0f), double((W + 1.0f) - W/2.0f), tx), lerp64(double(1.0f - H/2.0f), double((H + 1.0f) - H/2.0f), ty));
        fragCoord = stxy;
        instance = 0;
        gl_FragCoord = vec4(float(i) + 0.5f, float(j) + 0.5f, 0.5f, 1.0f);
        fragColor = vec4(0.0f);
        sfb_discarded = false;
        sfb_dscale = 1.0f;
    }

    // screen-space derivatives (GLSL 3.30 §8.13.1): differences inside the 2 x 2 quad a group of four lanes shades.
    // Undefined in non-uniform control flow, as in GLSL (lanes that are not executing contribute their own value).
    G_DEV float sfb_quad_difference(float v, int axis_bit) const {
        const float other = __shfl_xor_sync(__activemask(), v, axis_bit);
        return (((threadIdx.x & axis_bit) != 0) ? (v - other) : (other - v))*sfb_dscale;
    }
    template <class V> G_DEV V sfb_quad_difference_of(const V& v, int axis_bit) const {
        if constexpr (info<V>::n == 1) return V(sfb_quad_difference(float(v), axis_bit));
        else { V r; for (int k = 0; k < info<V>::n; k++) r.v[k] = sfb_quad_difference(v.v[k], axis_bit); return r; }
    }
    template <class V> G_DEV auto dFdx(const V& v) const { return sfb_quad_difference_of(v, 1); }
    template <class V> G_DEV auto dFdy(const V& v) const { return sfb_quad_difference_of(v, 2); }
    template <class V> G_DEV auto fwidth(const V& v) const { return abs(dFdx(v)) + abs(dFdy(v)); }
    G_DEV sampler2D sfb_sampler(int slot) const { sampler2D s; s.s = &sfb_params->tex[slot]; return s; }
    // module / user uniforms: float components as floats, int / uint / bool components as their 32 bits
    template <class T> G_DEV T sfb_extra(int slot) const {
        const float* e = sfb_params->u.extra[slot];
        using B = typename info<T>::base;
        if constexpr (B(0.5) != B(0)) {                          // float
            if constexpr (info<T>::n == 1) return T(e[0]);
            else return T(vec4(e[0], e[1], e[2], e[3]));
        } else {
            if constexpr (info<T>::n == 1) return T(__float_as_int(e[0]));
            else return T(ivec4(__float_as_int(e[0]), __float_as_int(e[1]), __float_as_int(e[2]), __float_as_int(e[3])));
        }
    }

    // a square matrix uniform: one slot per column
    template <int N> G_DEV mat<N, N> sfb_extra_matrix(int slot) const {
        mat<N, N> m;
        for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) m.c[j].v[i] = sfb_params->u.extra[slot + j][i];
        return m;
    }

    G_DEV vec2 agluv2gluv(vec2 p) const { return p*vec2(iAspectRatio, 1); }                                        // :99-100
    G_DEV vec2 gluv2agluv(vec2 p) const { return p/vec2(iAspectRatio, 1); }
    G_DEV vec2 stuv2stxy(vec2 s) const { return g::stuv2stxy(s, iResolution); }                                    // :104
    G_DEV vec2 stuv2stxy(vec2 s, vec2 r) const { return g::stuv2stxy(s, r); }
    G_DEV vec2 stxy2stuv(vec2 s) const { return g::stxy2stuv(s, iResolution); }                                    // :108
    G_DEV vec2 stxy2stuv(vec2 s, vec2 r) const { return g::stxy2stuv(s, r); }
    G_DEV vec2 astuv2stuv(vec2 s) const { return vec2(s.x*iAspectRatio + (1 - iAspectRatio)/2, s.y); }             // :111-116
    G_DEV vec2 stuv2astuv(vec2 s) const { return vec2((s.x - (1 - iAspectRatio)/2)/iAspectRatio, s.y); }
    G_DEV vec2 gluv_mirrored_repeat(vec2 p) const { return vec2(iWantAspect*triangle_wave(p.x, 4*iWantAspect), triangle_wave(p.y, 4.0f)); }   // :127
    G_DEV bool stuv_oob(vec2 s) const { return astuv_oob(stuv2astuv(s)); }                                         // :138
    G_DEV bool gluv_oob(vec2 p) const { return agluv_oob(gluv2agluv(p)); }                                         // :144
    G_DEV vec4 gtexture(sampler2D image, vec2 p) const { return g::gtexture(image, p); }
    G_DEV vec4 gmtexture(sampler2D image, vec2 p) const { return g::gtexture(image, gluv_mirrored_repeat(p)); }    // :173
    G_DEV vec4 gtexture(sampler2D image, vec2 p, bool mirror) const { return mirror ? gmtexture(image, p) : g::gtexture(image, p); }   // :178
    G_DEV vec4 agtexture(sampler2D image, vec2 p) const { return g::gtexture(image, agluv2gluv(p)); }              // :185
    G_DEV vec4 agmtexture(sampler2D image, vec2 p) const { return agtexture(image, agluv_mirrored_repeat(p)); }    // :191
    G_DEV vec4 agtexture(sampler2D image, vec2 p, bool mirror) const { return mirror ? agmtexture(image, p) : agtexture(image, p); }   // :196

    // camera.glsl:55-155
    G_DEV vec3 CameraRectangle(const Camera& c, vec2 p, float size) const { return size*(p.x*c.right + p.y*c.up); }
    G_DEV vec3 CameraRayOrigin(const Camera& c, vec2 p) const {
        return c.position + CameraRectangle(c, p, c.zoom*c.isometric) + (c.backward*c.orbital) + (c.backward*c.dolly); }
    G_DEV vec3 CameraRayTarget(const Camera& c, vec2 p) const {
        return c.position + CameraRectangle(c, p, c.zoom) + (c.backward*c.orbital) + (c.forward*c.focal_length); }
    G_DEV Camera CameraRay2D(Camera c) const {
        const float num = dot(c.plane_point - c.origin, c.plane_normal), den = dot(c.target - c.origin, c.plane_normal);
        const float t = num/den;
        c.out_of_bounds = (t < 0) || (abs(gluv.x) > iWantAspect);
        c.gluv = swz<0, 1>(c.origin + (t*(c.target - c.origin)));
        c.agluv = c.gluv/vec2(iAspectRatio, 1);
        c.stuv = (c.gluv + 1.0f)/2.0f;
        c.astuv = (c.agluv + 1.0f)/2.0f;
        c.stxy = iResolution*c.astuv;
        c.glxy = c.stxy - iResolution/2.0f;
        return c;
    }
    G_DEV Camera CameraProject(Camera c) const {
        if (c.projection == CameraProjectionPerspective) {
            c.origin = CameraRayOrigin(c, gluv);
            c.target = CameraRayTarget(c, gluv);
        } else if (c.projection == CameraProjectionStereoscopic) {
            const vec2 p = gluv - sign(agluv.x)*vec2(iAspectRatio/2.0f, 0.0f);
            c.position += (sign(agluv.x)*c.separation)*c.right;
            c.origin = CameraRayOrigin(c, p);
            c.target = CameraRayTarget(c, p);
        } else if (c.projection == CameraProjectionEquirectangular) {
            const float inclination = c.zoom*(PI*agluv.y/2), azimuth = c.zoom*(PI*agluv.x/1);
            vec3 target = c.forward;
            target = rotate3d(target, c.right, -inclination);
            target = rotate3d(target, c.up, azimuth);
            c.origin = c.position;
            c.target = c.position + target;
        }
        return CameraRay2D(c);
    }
    G_DEV Camera sfb_get_camera() const {            // the GetCamera(iCamera) macro
        Camera c;
        c.plane_point = vec3(0, 0, 1); c.plane_normal = vec3(0, 0, 1);
        c.mode = iCameraMode; c.projection = iCameraProjection; c.position = iCameraPosition;
        c.orbital = iCameraOrbital; c.dolly = iCameraDolly; c.zenith = iCameraZenith;
        c.up = iCameraUpward; c.down = iCameraUpward*(-1); c.left = iCameraRight*(-1); c.right = iCameraRight;
        c.forward = iCameraForward; c.backward = iCameraForward*(-1);
        c.isometric = iCameraIsometric; c.focal_length = iCameraFocalLength; c.zoom = iCameraZoom; c.separation = iCameraSeparation;
        c.out_of_bounds = false;
        return CameraProject(c);
    }
};

}  // namespace g
