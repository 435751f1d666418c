// Calls ShaderFlow's GLSL std-lib API (shaderflow.glsl, camera.glsl) group by group; `iProbe` selects the group.
// The golden (tests/golden/jit_stdlib.npz) is this text evaluated behind the REFERENCE's header and include files;
// the CUDA backend compiles it against its own implementation of the same API (csrc/jit/shaderflow_rt.cuh).
uniform int iProbe;

vec4 probe_camera() {
    GetCamera(iCamera);
    if (iCamera.out_of_bounds) return vec4(0.25);
    return vec4(iCamera.gluv, iCamera.astuv.x, dot(iCamera.target - iCamera.origin, vec3(0.1, 0.2, 0.3)) + iCamera.stxy.y/iHeight);
}

vec4 probe_coordinates() {
    vec2 a = stuv2gluv(astuv) + s2g(stuv)*0.01;
    vec2 b = gluv2stuv(gluv) + g2s(agluv)*0.01;
    vec2 c = agluv2gluv(agluv) - gluv2agluv(gluv);
    vec2 d = stuv2stxy(astuv)/iResolution + stxy2stuv(stxy) + stuv2stxy(astuv, vec2(3, 5)) - stxy2stuv(stxy, vec2(7, 9));
    vec2 e = astuv2stuv(astuv) - stuv2astuv(stuv);
    float oob = float(astuv_oob(astuv*1.5 - 0.2)) + 2*float(stuv_oob(stuv*1.2)) + 4*float(agluv_oob(agluv*1.3)) + 8*float(gluv_oob(gluv*1.1));
    return vec4(a.x + b.y + c.x, d.x - d.y, e.x + e.y + iAspectRatio + iCycle + iFrametime + iDeltatime, oob/16 + iWidth/iHeight);
}

vec4 probe_interpolation() {
    float x = agluv.x, y = agluv.y;
    float a = proportion(2.0, x + 3.0, y + 2.0) + lerp(-1.0, 2.0, 1.0, -3.0, x);
    float b = smoothlerp(x, y, 0.3) + smin(x, y, 0.2) + smax(x, y, 0.4) + smin(x, y) - smax(x, y);
    float c = smoothmix(1.0, 3.0, -0.5, 0.5, x) + smix(2.0, -1.0, -0.2, 0.8, y);
    float d = triangle_wave(x*3.0, 1.5) + triangle_wave(y, 0.7);
    return vec4(a, b, c, d);
}

vec4 probe_angles() {
    vec3 v = normalize(vec3(gluv, 1.0));
    float a = angle(gluv + vec2(2, 0.5), vec2(1, 0.3)) + angle(v, vec3(0, 1, 0)) + angle(vec4(v, 1), vec4(1, 2, 3, 4));
    vec2 r = rotate2d(0.7 + gluv.x)*gluv + rotate2deg(33.0)*agluv;
    vec3 q = rotate3d(v, normalize(vec3(1, 2, 3)), gluv.y) + rotate3deg(v, vec3(0, 0, 1), 75.0);
    float t = atan_normalized(gluv.x*3) + atan1(gluv + 0.01) + atan1n(gluv + 0.01) + atan2(gluv.y + 0.01, gluv.x) + atan2(gluv + 0.02)
            + atan2n(gluv.y + 0.01, gluv.x) + atan2n(gluv + 0.03);
    vec2 p = polar2rect(1.0 + agluv.x, gluv.y*3.0);
    vec3 s = sphere2rect(2.0, agluv.x*3.0, agluv.y*2.0);
    return vec4(a + r.x, r.y + q.x + q.y*0.5 + q.z*0.25, t, p.x - p.y + s.x + s.y*0.5 + s.z*0.25);
}

vec4 probe_color() {
    vec3 a = palette_magma(astuv.x) + palette(astuv.y, vec3(0.1), vec3(0.9, 0.2, 0.1), vec3(0.2, 0.8, 0.3), vec3(0.1, 0.3, 1.0));
    vec3 h = hsv2rgb(vec3(astuv.x*7.0, 0.8, 0.9)) + hsv2rgb(astuv.y*6.0, 0.5, 1.0) + hsv2rgb(vec4(1.0, astuv, 0.5)).rgb;
    vec3 k = rgb2hsv(vec3(astuv, 0.3)) + rgb2hsv(0.2, astuv.x, astuv.y) + rgb2hsv(vec4(astuv.yx, 0.9, 1.0)).xyz;
    vec4 b = blend(vec4(a, 0.5), vec4(h, astuv.x)) + alpha_composite(vec4(k, 1.0), vec4(a, astuv.y));
    vec4 s = saturate(b, 0.3) + vec4(saturate(h, 0.4), 0) + vec4(saturate(astuv, 1.7), 0, 0);
    float keys = float(isBlackKey(int(stxy.x))) + 2*float(isWhiteKey(stxy.y)) + 4*float(isBlackKey(stxy.y + 3.0)) + 8*float(isWhiteKey(int(stxy.x) + 5));
    return s + vec4(0, 0, 0, keys/16);
}

vec4 probe_distance() {
    vec3 o = vec3(gluv*2.0, 0.5);
    float a = sdLine(gluv, vec2(-1, -0.5), vec2(1, 0.7)) + sdLine(o, vec3(0), vec3(1, 1, 1)) + sdLineSegment(o, vec3(-1, 0, 0), vec3(0.5, 0.5, 1))
            + sdLineSegment(gluv, vec2(-0.3, 0.2), vec2(0.4, -0.6));
    float b = sdSphere(o, vec3(0.3, 0.2, 1.0), 0.8) + sdPlane(o, vec3(0, -1, 0), vec3(0.1, 1, 0.2)) + sdBox(o, vec3(0.2, 0, 1), vec3(1, 0.6, 0.8))
            + sdOctahedron(o, vec3(-0.5, 0.1, 0.7), 0.9);
    float x = gluv.x, y = gluv.y;
    float c = sdUnion(x, y) + sdSmoothUnion(x, y, 0.3) + sdSubtraction(x, y) + sdSmoothSubtraction(x, y, 0.25);
    float d = sdIntersection(x, y) + sdSmoothIntersection(x, y, 0.35);
    return vec4(a, b, c, d);
}

vec4 probe_textures() {
    vec2 z = zoom(gluv, 0.8 + 0.1*sin(iTime), vec2(0.2, -0.1)) + zoom(agluv, 1.1)*0.01;
    vec4 a = gtexture(background, z) + stexture(background, stuv) + astexture(background, astuv);
    vec4 b = agtexture(background, agluv*1.7) + gmtexture(background, gluv*1.9) + agmtexture(background, agluv*2.3);
    vec4 c = gtexture(background, gluv*2.1, true) + gtexture(background, gluv*0.7, false) + agtexture(background, agluv*2.6, true)
           + agtexture(background, agluv*0.4, false);
    vec2 m = agluv_mirrored_repeat(agluv*3.0) + gluv_mirrored_repeat(gluv*2.5);
    return (a + b + c)/10.0 + vec4(m, 0, 0)*0.1;
}

void main() {
    switch (iProbe) {
        case 0: fragColor = probe_camera(); break;
        case 1: fragColor = probe_coordinates(); break;
        case 2: fragColor = probe_interpolation(); break;
        case 3: fragColor = probe_angles(); break;
        case 4: fragColor = probe_color(); break;
        case 5: fragColor = probe_distance(); break;
        default: fragColor = probe_textures();
    }
}
